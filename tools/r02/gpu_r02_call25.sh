#!/bin/bash
# L2-resident image groups for the generator chain: same-box A/B
for g in 0 2 4 8 16 0; do
  timeout 200 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --l2-group $g 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('l2_group=$g', round(d['ms_per_step'], 3), 'ms', round(d['value']), 'fps', d['gpu_launches'])"
done
