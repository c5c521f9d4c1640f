#!/bin/sh
KREGEX='regex:scan_kernel|prefix_sum_kernel|prefix_scan_kernel|place_kernel|redo_kernel|stats|no_overlap|transform|window|visible|fold_'
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2e_tests.log
for w in names-cpw names cfg4; do echo "== $w"; timeout 300 python tools/profile_scan.py --size-gib 4 --workload $w --iters 3 2>&1 | tail -1; done
python tools/profile_scan.py --size-gib 1 --workload names-cpw --iters 2 > gpurun_out/r2e_p_cpw.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 120 --csv --log-file gpurun_out/r2e_launches_cpw.csv python tools/profile_scan.py --size-gib 1 --workload names-cpw --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^scan_kernel" -s 4 -c 1 -f -o gpurun_out/r2e_scan_cpw python tools/profile_scan.py --size-gib 1 --workload names-cpw --iters 2 > gpurun_out/r2e_ncu_cpw.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^scan_kernel" -s 1 -c 1 -f -o gpurun_out/r2e_scan_names python tools/profile_scan.py --size-gib 2 --workload names --iters 2 > gpurun_out/r2e_ncu_names.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^scan_kernel" -s 1 -c 1 -f -o gpurun_out/r2e_scan_cfg4 python tools/profile_scan.py --size-gib 2 --workload cfg4 --iters 2 > gpurun_out/r2e_ncu_cfg4.log 2>&1
echo "ncu rc=$?"
