#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for fpl in 8 11 13 16 22 32 0; do
  python bench.py --steps 10 --warmup 3 --no-extras --repeats 3 --frames-per-launch $fpl 2>/dev/null | python -c "
import sys, json
j=json.loads([x for x in sys.stdin if x.startswith('{')][-1])
print('fpl', $fpl, 'value', j['value'], 'e2e', j['e2e']['value'], j['roofline']['kernel_ms_per_launch'], j['impl_config'])"
done
