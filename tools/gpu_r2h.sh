#!/bin/sh
# profiles of the census legs (BASELINE configs[1], [2]) + launch lists
KREGEX='regex:scan_kernel|prefix_sum_kernel|prefix_scan_kernel|place_kernel|redo_kernel|stats|no_overlap|transform|window|visible|fold_'
for w in census-c census-cpw; do
  python bench.py --leg $w --no-cpu > gpurun_out/r2h_leg_$w.json 2> gpurun_out/r2h_leg_$w.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 100 --csv --log-file gpurun_out/r2h_launches_$w.csv python bench.py --leg $w --no-cpu > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k "regex:^scan_kernel" -s 2 -c 1 -f -o gpurun_out/r2h_scan_$w python bench.py --leg $w --no-cpu > gpurun_out/r2h_ncu_$w.log 2>&1
  echo "$w rc=$?"
done
