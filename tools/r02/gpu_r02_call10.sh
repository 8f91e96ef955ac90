#!/bin/bash
# round 2, GPU call 10: pixel-shuffle epilogue of the CTA-pair general conv (F = 128 upsampling); BASELINE configs[4] sweep.
O=gpurun_out/r02
mkdir -p $O
timeout 300 python -m pytest tests/test_gen_2cta_gpu.py tests/test_generator_gpu.py -q -x > $O/t_call10.log 2>&1; echo "rc=$?" >> $O/t_call10.log; tail -3 $O/t_call10.log
timeout 600 python tools/sweep.py > $O/sweep_generator_filters_layers_sizes.md 2> $O/sweep.err; cat $O/sweep_generator_filters_layers_sizes.md
