#!/bin/sh
# short candidates' second look in shared memory (sx): parity suite, then synthetic workloads and the text legs next to the previous library
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2w_tests.log
WL="cfg4 names names-cpw" sh tools/gpu_variants.sh > gpurun_out/r2w_variants.log 2>&1; cat gpurun_out/r2w_variants.log
for lib in omega_match_b200/lib/libomega_match.so omega_match_b200/lib/variants/v_prev.so; do
  for leg in names census-c census-cpw cfg4; do
    OMEGA_MATCH_LIB_PATH=$PWD/$lib python bench.py --leg $leg --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib'.split('/')[-1], '$leg', round(d['value'], 1), 'GB/s', d.get('matches_per_step'))"
  done
done
