#!/bin/bash
# weight-stationary general conv: parity first, then the train step with it on / off, launch list, full suite, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gen_ws_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/t_genws.log
tail -3 gpurun_out/t_genws.log
for ws in 1 0; do
  FSR_GEN_WS=$ws timeout 200 python tools/bench_train.py --batch 64 --steps 20 --warmup 5 > gpurun_out/train_genws$ws.json 2> gpurun_out/train_genws$ws.err
  cat gpurun_out/train_genws$ws.json
done
timeout 420 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -15 > gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 1550 -c 520 --csv --log-file gpurun_out/launches_trainstep_b64.csv \
    python tools/bench_train.py --batch 64 --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_gen_ws" -s 20 -c 6 \
    -o gpurun_out/genws_full python tools/bench_train.py --batch 64 --steps 1 --warmup 1 > gpurun_out/ncu_genws.log 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
head -c 300 gpurun_out/bench_n1.json
