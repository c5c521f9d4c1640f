#!/bin/sh
# ncu capture of the small-store kernel on the short-pattern store (configs[3]) with the second look on
ncu --set full --clock-control none --import-source on -k "regex:^scan_kernel" -s 1 -c 1 -f -o gpurun_out/r2x_scan_cfg4 python tools/profile_scan.py --size-gib 2 --workload cfg4 --iters 2 > gpurun_out/r2x_ncu_cfg4.log 2>&1
echo "ncu rc=$?"
