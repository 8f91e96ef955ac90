#!/bin/bash
# programmatic dependent launch: correctness (train-step + determinism tests) and same-box A/B
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_train_ops_gpu.py tests/test_generator_gpu.py tests/test_baseline_configs_gpu.py -x -q -m gpu 2>&1 | tail -4
for pdl in 0 1 0 1; do
  FSR_PDL=$pdl timeout 300 python tools/bench_train.py --batch 64 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl=$pdl b64', round(d['ms_per_step'], 3), 'ms', d['losses'])"
done
for pdl in 0 1; do
  FSR_PDL=$pdl timeout 300 python tools/bench_train.py --batch 32 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl=$pdl b32', round(d['ms_per_step'], 3), 'ms')"
  FSR_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl=$pdl gen', round(d['ms_per_step'], 3), 'ms', round(d['value']), 'fps')"
done
