#!/bin/bash
# profiles/stage_times.sh [lib.so ...] -- device-resident leg of bench.py per library build (kernel experiments)
for lib in "$@"; do
  WN_B200_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu --no-pbc --no-find 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lib', 'fps %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], {k: round(v,3) for k,v in r['stage_ms_per_step'].items()})"
done
