#!/bin/bash
# full GPU suite + headline bench after the pair-row changes (main kernels' MMA issue loop was touched)
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_after_pairs.json 2> gpurun_out/r02/bench_after_pairs.err
cat gpurun_out/r02/bench_after_pairs.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'roofline', 'clocks')})
print(d.get('train_step', {}).get('b64', d.get('train_step')))
"
