#!/bin/bash
# fused-input conv, iteration 2 (4 transform warps, branch-free): parity, A/B bench, full suite in both modes, capture
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fuse_in_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/t_fuse.log
tail -2 gpurun_out/t_fuse.log
FSR_FUSE_IN=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/bench_n1_fuse1.json 2> gpurun_out/bench_fuse1.err
head -c 250 gpurun_out/bench_n1_fuse1.json; echo
FSR_FUSE_IN=0 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1_fuse0.json 2> gpurun_out/bench_fuse0.err
head -c 250 gpurun_out/bench_n1_fuse0.json; echo
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/t_all_default.log
tail -2 gpurun_out/t_all_default.log
FSR_FUSE_IN=1 timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/t_all_fuse1.log
tail -2 gpurun_out/t_all_fuse1.log
FSR_FUSE_IN=1 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_generator_fuse1.csv \
    python tools/profile_step.py 2 > gpurun_out/ncu_gen.log 2>&1
FSR_FUSE_IN=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_c64_kernel" -s 24 -c 2 \
    -o gpurun_out/resblock_xf2_full python tools/profile_step.py 2 > gpurun_out/ncu_xf.log 2>&1
tail -1 gpurun_out/ncu_xf.log
