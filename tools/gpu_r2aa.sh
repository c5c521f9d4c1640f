#!/bin/sh
# ncu capture of one scan launch of the census-c leg (BASELINE configs[1]: ignore-case + word_boundary; the generic per-candidate path)
ncu --set full --clock-control none --import-source on -k "regex:^scan_kernel" -s 3 -c 1 -f -o gpurun_out/r2aa_scan_census-c python bench.py --leg census-c --no-cpu > gpurun_out/r2aa_ncu.log 2>&1
echo "ncu rc=$?"
