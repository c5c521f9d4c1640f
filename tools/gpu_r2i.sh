#!/bin/sh
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2i_tests.log
for w in cfg5 names-synth cfg4; do echo "== $w"; timeout 300 python tools/profile_scan.py --size-gib 4 --workload $(echo $w | sed 's/names-synth/names/') --iters 3 2>&1 | tail -1; done
for w in names census-c census-cpw; do echo "== leg $w"; timeout 300 python bench.py --leg $w --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['matches_per_step'], d['roofline']['scan_ms'], d['roofline']['filter_ms'])"; done
