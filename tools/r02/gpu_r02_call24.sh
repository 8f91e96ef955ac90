#!/bin/bash
# XF stat decode spread over lanes; fused IN backward with shared stat decode: parity + timings
timeout 900 python -m pytest tests/test_fuse_in_gpu.py tests/test_generator_gpu.py tests/test_train_step_gpu.py tests/test_pairs_gpu.py tests/test_fused_chain_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
timeout 300 python tools/bench_train.py --batch 64 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b64', round(d['ms_per_step'], 3), 'ms', d['losses']['loss_real'])"
done
timeout 300 python tools/bench_train.py --batch 32 --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b32', round(d['ms_per_step'], 3), 'ms')"
timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gen', round(d['ms_per_step'], 3), 'ms', round(d['value']), 'fps', d['train_step'].get('b64', {}).get('ms_per_step'), d['clocks'])"
