#!/bin/bash
# one gpurun call: new-kernel tests, micro-benchmarks, headline bench (A/B small_mma), full GPU suite
mkdir -p gpurun_out
python -m pytest tests/test_small_mma_gpu.py tests/test_aux_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/t_new.log
python tools/bench_aux.py > gpurun_out/bench_aux.json 2> gpurun_out/bench_aux.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
FSR_SMALL_MMA=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_smallmma0.json 2> /dev/null
timeout 400 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -30 > gpurun_out/t_all.log
tail -5 gpurun_out/t_new.log; cat gpurun_out/bench_aux.json; tail -3 gpurun_out/t_all.log
