#!/bin/bash
# stream-balance A/B: where does the D(sr) adversarial branch run once the whole D step is on the side stream?
for cfg in "1 1" "1 0" "1 1" "1 0"; do
  set -- $cfg
  FSR_D_SIDE=$1 FSR_ADV_SIDE=$2 timeout 300 python tools/bench_train.py --batch 64 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('d_side=$1 adv_side=$2 b64', round(d['ms_per_step'], 3), 'ms')"
done
for cfg in "1 1" "1 0"; do
  set -- $cfg
  FSR_D_SIDE=$1 FSR_ADV_SIDE=$2 timeout 300 python tools/bench_train.py --batch 32 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('d_side=$1 adv_side=$2 b32', round(d['ms_per_step'], 3), 'ms')"
done
