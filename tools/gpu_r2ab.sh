#!/bin/sh
# the lean path with start predicates (flagged calls leave the generic path): parity suite, then the census-c leg (word_boundary) and cfg5 with word_boundary
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2ab_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ab_tests.log
for leg in census-c census-cpw; do
  python bench.py --leg $leg --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$leg', round(d['value'], 1), 'GB/s', d.get('matches_per_step'))"
done
python tools/profile_scan.py --size-gib 4 --workload cfg5 --iters 3 --flags word_boundary 2>&1 | tail -1
python tools/profile_scan.py --size-gib 4 --workload names --iters 3 --flags word_boundary 2>&1 | tail -1
