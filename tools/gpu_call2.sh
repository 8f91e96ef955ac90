#!/bin/bash
# full GPU suite, micro-benchmarks, ncu launch lists + full captures of the small kernels, headline bench, incumbent
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -30 > gpurun_out/t_all.log
python tools/bench_aux.py > gpurun_out/bench_aux.json 2> gpurun_out/bench_aux.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"neck_conv3x3_mma|wgrad_c3_mma|psnr_ssim|crop_resize" -c 10 \
    -o gpurun_out/aux_full python tools/profile_aux.py > gpurun_out/ncu_aux.log 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 45 --csv --log-file gpurun_out/launches_generator_b32_180x320.csv \
    python tools/profile_step.py 2 > gpurun_out/ncu_gen.log 2>&1
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 1550 -c 520 --csv --log-file gpurun_out/launches_trainstep_b64.csv \
    python tools/bench_train.py --batch 64 --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 200 python tools/incumbent.py --steps 10 --no-compile > gpurun_out/incumbent.json 2> gpurun_out/incumbent.err
tail -4 gpurun_out/t_all.log; cat gpurun_out/bench_aux.json; head -c 400 gpurun_out/bench_n1.json; echo; cat gpurun_out/incumbent.json; tail -2 gpurun_out/ncu_aux.log
