#!/bin/bash
# deterministic small reductions: op parity, train-step tests, cross-process reproducibility, timing
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_train_ops_gpu.py tests/test_train_step_gpu.py tests/test_autograd_bridge_gpu.py tests/test_baseline_configs_gpu.py tests/test_small_mma_gpu.py -x -q -m gpu 2>&1 | tail -6
for i in 1 2; do
  timeout 300 python tools/bench_train.py --batch 64 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run $i b64', round(d['ms_per_step'], 3), 'ms', d['losses'])"
done
timeout 300 python tools/bench_train.py --batch 32 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b32', round(d['ms_per_step'], 3), 'ms')"
