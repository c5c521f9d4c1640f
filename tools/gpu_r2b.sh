#!/bin/sh
# round 2, second GPU call: new config-scale tests, grab/tile variants at cfg5, ncu of the small-store kernels
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2b_tests.log
WL=cfg5 sh tools/gpu_variants.sh > gpurun_out/r2b_variants.log 2>&1; cat gpurun_out/r2b_variants.log
for w in names cfg4; do
  python tools/profile_scan.py --size-gib 4 --workload $w --iters 2 > gpurun_out/r2b_p_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:::scan_kernel" -c 1 -f -o gpurun_out/r2b_scan_$w \
      python tools/profile_scan.py --size-gib 4 --workload $w --iters 2 > gpurun_out/r2b_ncu_$w.log 2>&1
  echo "$w ncu rc=$?"
done
