#!/bin/sh
# flagged calls: the lean path with start predicates (default library) next to the generic path (v_prev)
for lib in omega_match_b200/lib/libomega_match.so omega_match_b200/lib/variants/v_prev.so; do
  echo "== $lib"
  for w in cfg5 names cfg4; do
    OMEGA_MATCH_LIB_PATH=$PWD/$lib python tools/profile_scan.py --size-gib 4 --workload $w --iters 3 --flags word_boundary 2>&1 | tail -1
  done
  OMEGA_MATCH_LIB_PATH=$PWD/$lib python tools/profile_scan.py --size-gib 4 --workload cfg5 --iters 3 --flags line_start,longest_only 2>&1 | tail -1
done
