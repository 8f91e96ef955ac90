#!/bin/bash
# fused InstanceNorm+PReLU input transform of the res-block conv: parity, A/B bench, launch list, full capture, full suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fuse_in_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/t_fuse.log
tail -3 gpurun_out/t_fuse.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
head -c 250 gpurun_out/bench_n1.json; echo
FSR_FUSE_IN=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/bench_n1_fuse0.json 2> /dev/null
head -c 250 gpurun_out/bench_n1_fuse0.json; echo
timeout 420 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -12 > gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_generator_b32_180x320.csv \
    python tools/profile_step.py 2 > gpurun_out/ncu_gen.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_c64_kernel" -s 24 -c 4 \
    -o gpurun_out/resblock_xf_full python tools/profile_step.py 2 > gpurun_out/ncu_xf.log 2>&1
tail -2 gpurun_out/ncu_xf.log
