#!/bin/bash
# register-resident fused InstanceNorm backward: tests + same-box A/B
timeout 900 python -m pytest tests/test_train_ops_gpu.py tests/test_train_step_gpu.py tests/test_baseline_configs_gpu.py tests/test_autograd_bridge_gpu.py -x -q -m gpu 2>&1 | tail -4
for d in 0 1 0 1; do
  FSR_IN_BWD_RES=$d timeout 300 python tools/bench_train.py --batch 64 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('in_bwd_res=$d b64', round(d['ms_per_step'], 3), 'ms', d['losses']['loss_real'], d['losses']['content_loss'])"
done
for d in 0 1; do
  FSR_IN_BWD_RES=$d timeout 300 python tools/bench_train.py --batch 32 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('in_bwd_res=$d b32', round(d['ms_per_step'], 3), 'ms')"
done
