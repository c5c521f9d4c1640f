#!/bin/sh
# the second look forced on for names.txt (hashed three-byte p23): synthetic haystack and the KJV-like leg
for v in 0 1; do
  export OLM_SHORT_LOOK=$v
  echo "OLM_SHORT_LOOK=$v"
  python tools/profile_scan.py --size-gib 4 --workload names --iters 3 2>&1 | tail -1
  python bench.py --leg names --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('names leg', round(d['value'], 1), 'GB/s', d.get('matches_per_step'))"
done
