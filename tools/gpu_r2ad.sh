#!/bin/sh
# threshold of the warp-cooperative record compare (records behind a key): 8 (default) vs 4 vs 16 on the census legs
for lib in omega_match_b200/lib/libomega_match.so omega_match_b200/lib/variants/v_coop4.so omega_match_b200/lib/variants/v_coop16.so; do
  for leg in census-c census-cpw; do
    OMEGA_MATCH_LIB_PATH=$PWD/$lib python bench.py --leg $leg --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib'.split('/')[-1], '$leg', round(d['value'], 1), 'GB/s', d.get('matches_per_step'))"
  done
done
