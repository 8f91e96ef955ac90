#!/bin/bash
# pixel-pair path for n_filters = 32: parity tests + sweep rows
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_pairs_gpu.py tests/test_generator_gpu.py -x -q -m gpu -s 2>&1 | tail -30
echo "rc=$?"
timeout 600 python tools/sweep.py 32 > gpurun_out/r02/sweep_f32_pairs.md 2> gpurun_out/r02/sweep_f32_pairs.err
cat gpurun_out/r02/sweep_f32_pairs.md
FSR_PAIR32=0 timeout 600 python tools/sweep.py 32 > gpurun_out/r02/sweep_f32_padded.md 2>> gpurun_out/r02/sweep_f32_pairs.err
cat gpurun_out/r02/sweep_f32_padded.md
tail -5 gpurun_out/r02/sweep_f32_pairs.err
