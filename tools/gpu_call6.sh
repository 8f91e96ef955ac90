#!/bin/bash
# last slot of the round: fp32 stat_fix + rolled XF transform: parity of the touched kernels, then A/B bench
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_fuse_in_gpu.py tests/test_ops_gpu.py tests/test_gen_ws_gpu.py tests/test_generator_gpu.py -q -x 2>&1 | tail -4 > gpurun_out/t_last.log
tail -2 gpurun_out/t_last.log
FSR_FUSE_IN=1 timeout 60 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench_last_fuse1.json 2> /dev/null
head -c 230 gpurun_out/bench_last_fuse1.json; echo
timeout 80 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last_fuse0.json 2> /dev/null
head -c 230 gpurun_out/bench_last_fuse0.json; echo
