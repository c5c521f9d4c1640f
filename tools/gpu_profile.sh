#!/bin/sh
# Round-end evidence on the GPU box (run through gpurun): plain bench line, ncu launch list of the
# same command, one full ncu capture of the scan kernel at the bench size.  $1 = tag (e.g. r1j)
T=${1:-rX}
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    -k "regex:scan_kernel|prefix_sum_kernel|prefix_scan_kernel|place_kernel|redo_kernel|stats_kernel|no_overlap|transform" --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-experimental > gpurun_out/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -c 1 -f -o gpurun_out/${T}_scan_cfg5_16g \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-experimental > gpurun_out/${T}_ncu_full.log 2>&1
cut -c1-400 gpurun_out/${T}_bench.json
