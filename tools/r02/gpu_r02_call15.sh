#!/bin/bash
# pixel-pair path with structural-zero K-step skipping: parity + sweep rows, skip on/off
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_pairs_gpu.py tests/test_generator_gpu.py tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/sweep.py 32 > gpurun_out/r02/sweep_f32_pairs_skip.md 2> gpurun_out/r02/sweep_f32_pairs.err
cat gpurun_out/r02/sweep_f32_pairs_skip.md
FSR_PAIR_SKIP=0 timeout 600 python tools/sweep.py 32 | tail -8
tail -3 gpurun_out/r02/sweep_f32_pairs.err
