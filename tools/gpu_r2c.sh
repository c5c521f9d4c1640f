#!/bin/sh
# quick parity of the new scan (private chunk buffers, fused normalisation), then throughput
timeout 300 python tests/gpu_quick.py > gpurun_out/r2c_quick.log 2>&1; echo "quick rc=$?"; grep -c "OK " gpurun_out/r2c_quick.log; grep "BAD" gpurun_out/r2c_quick.log | head -20; tail -3 gpurun_out/r2c_quick.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2c_tests.log
for w in cfg5 names cfg4 names-cpw; do
  echo "== $w"; timeout 300 python tools/profile_scan.py --size-gib 4 --workload $w --iters 3 2>&1 | tail -2
done
echo "== cfg5 OLM_PRIV=0"; OLM_PRIV=0 timeout 300 python tools/profile_scan.py --size-gib 4 --workload cfg5 --iters 3 2>&1 | tail -1
echo "== names OLM_PRIV=0"; OLM_PRIV=0 timeout 300 python tools/profile_scan.py --size-gib 4 --workload names --iters 3 2>&1 | tail -1
