#!/bin/sh
# round 2, first GPU call: probe microbenchmarks, span-path check, baseline profiles of the small-store kernels
timeout 180 tools/microbench/probe_bench > gpurun_out/r2a_probe_bench.log 2>&1
echo "probe_bench rc=$?"
timeout 600 python tests/gpu_span_check.py > gpurun_out/r2a_span.log 2>&1
echo "span rc=$?"; tail -3 gpurun_out/r2a_span.log
for w in names cfg4; do
  python tools/profile_scan.py --size-gib 4 --workload $w --iters 3 > gpurun_out/r2a_p_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o gpurun_out/r2a_scan_$w \
      python tools/profile_scan.py --size-gib 4 --workload $w --iters 3 > gpurun_out/r2a_ncu_$w.log 2>&1
  echo "$w rc=$?"; cat gpurun_out/r2a_p_$w.log | tail -3
done
python tools/profile_scan.py --size-gib 4 --workload names-cpw --iters 3 2>&1 | tail -3
