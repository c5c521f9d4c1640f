#!/bin/sh
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_multi.py -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2f_tests.log
for w in cfg5 names-cpw names cfg4; do echo "== $w"; timeout 300 python tools/profile_scan.py --size-gib 4 --workload $w --iters 3 2>&1 | tail -1; done
echo "== names-cpw word_boundary"; timeout 300 python tools/profile_scan.py --size-gib 4 --workload names-cpw --iters 3 --flags word_boundary 2>&1 | tail -1
