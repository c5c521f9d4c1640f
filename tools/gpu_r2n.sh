#!/bin/sh
# ncu captures of the per-warp-slot scan kernel: cfg5 and names at 2 GiB
ncu --set full --clock-control none --import-source on -k "regex:^scan_kernel" -s 1 -c 1 -f -o gpurun_out/r2n_scan_cfg5 python tools/profile_scan.py --size-gib 2 --workload cfg5 --iters 2 > gpurun_out/r2n_ncu_cfg5.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^scan_kernel" -s 1 -c 1 -f -o gpurun_out/r2n_scan_names python tools/profile_scan.py --size-gib 2 --workload names --iters 2 > gpurun_out/r2n_ncu_names.log 2>&1
echo "ncu rc=$?"
