#!/bin/bash
# One gpurun call that validates a change end to end (what every GPU call of round 1 ran, in this order):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_validate.sh'
# parity suite -> headline bench -> micro-benchmarks -> ncu launch lists (generator forward, GAN train step).
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -12 > gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
head -c 300 gpurun_out/bench_n1.json; echo
python tools/bench_aux.py > gpurun_out/bench_aux.json 2> gpurun_out/bench_aux.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 32 --csv --log-file gpurun_out/launches_generator_b32_180x320.csv \
    python tools/profile_step.py 2 > gpurun_out/ncu_gen.log 2>&1
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 1550 -c 520 --csv --log-file gpurun_out/launches_trainstep_b64.csv \
    python tools/bench_train.py --batch 64 --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
