#!/bin/sh
# GPU box: the experimental stride-2 sampled scan (OLM_SAMPLE2=1) --
# parity against the oracle first, then throughput next to the default path.
#   gpurun --timeout 900 -- 'sh tools/gpu_s2.sh > gpurun_out/s2.log 2>&1'
set -x
OLM_SAMPLE2=1 timeout 600 python tests/gpu_s2_check.py | tail -60 || echo "S2 PARITY FAILED"
for w in cfg5 names; do
  for s2 in 0 1; do
    echo "== workload $w OLM_SAMPLE2=$s2"
    OLM_SAMPLE2=$s2 timeout 300 python tools/profile_scan.py --size-gib 4 --workload $w --iters 3 2>&1 | tail -3
  done
done
# multi-flag sanity on the big store: same counts with and without the mode
for s2 in 0 1; do
  OLM_SAMPLE2=$s2 timeout 300 python tools/profile_scan.py --size-gib 1 --workload cfg5 --iters 1 --flags longest_only,no_overlap 2>&1 | tail -1
done
# variant builds that travel with the snapshot (omega_match_b200/build/ does not): build them HERE first, e.g.
#   make -C omega_match_b200/csrc -j8 BUILD=../build/v_grab2 OUT=../lib/v_grab2/libomega_match.so EXTRA=-DOLM_GRAB=2
for lib in omega_match_b200/lib/v_*/libomega_match.so; do
  [ -f "$lib" ] || continue
  for s2 in 0 1; do
    echo "== $lib OLM_SAMPLE2=$s2"
    OMEGA_MATCH_LIB_PATH=$PWD/$lib OLM_SAMPLE2=$s2 timeout 300 python tools/profile_scan.py --size-gib 4 --workload cfg5 --iters 3 2>&1 | tail -1
  done
done
