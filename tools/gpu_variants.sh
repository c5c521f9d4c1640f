#!/bin/sh
# perf of variant builds (omega_match_b200/lib/variants/v_*.so) next to the default library
for lib in omega_match_b200/lib/libomega_match.so omega_match_b200/lib/variants/v_*.so; do
  echo "== $lib"
  for w in ${WL:-cfg5 cfg4 names}; do
    OMEGA_MATCH_LIB_PATH=$PWD/$lib python tools/profile_scan.py --size-gib 4 --workload $w --iters 3 2>&1 | tail -1
  done
done
