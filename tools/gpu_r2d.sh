#!/bin/sh
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2d_tests.log
for w in cfg5; do for pv in 1 0; do echo "== $w OLM_PRIV=$pv"; OLM_PRIV=$pv timeout 300 python tools/profile_scan.py --size-gib 4 --workload $w --iters 3 2>&1 | tail -1; done; done
python tools/profile_scan.py --size-gib 1 --workload names-cpw --iters 2 > gpurun_out/r2d_p_cpw.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2d_launches_cpw.csv python tools/profile_scan.py --size-gib 1 --workload names-cpw --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^scan_kernel" -s 4 -c 1 -f -o gpurun_out/r2d_scan_cpw python tools/profile_scan.py --size-gib 1 --workload names-cpw --iters 2 > gpurun_out/r2d_ncu_cpw.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2d_p_cpw.log
