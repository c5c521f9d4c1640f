#!/bin/sh
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_multi.py -x -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2j_tests.log
for w in cfg5 names cfg4 names-cpw; do echo "== $w"; timeout 300 python tools/profile_scan.py --size-gib 4 --workload $w --iters 3 2>&1 | tail -1; done
for w in names census-c census-cpw; do echo "== leg $w"; timeout 300 python bench.py --leg $w --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['matches_per_step'], d['roofline']['scan_ms'], d['roofline']['filter_ms'])"; done
ncu --set full --clock-control none --import-source on -k "regex:^scan_kernel" -s 2 -c 1 -f -o gpurun_out/r2j_scan_census-cpw python bench.py --leg census-cpw --no-cpu > gpurun_out/r2j_ncu.log 2>&1
