#!/bin/bash
# round 2, GPU call 4: train step - [sr;hr] batch, grouped deterministic wgrad, fused InstanceNorm backward, stats arena.
O=gpurun_out/r02
mkdir -p $O
timeout 600 python -m pytest tests/test_train_ops_gpu.py tests/test_train_step_gpu.py tests/test_baseline_configs_gpu.py tests/test_gen_ws_gpu.py -q -x -k "not generator_180 and not checkpoint" > $O/t_call4.log 2>&1; echo "rc=$?" >> $O/t_call4.log
tail -5 $O/t_call4.log
for B in 64 32; do
  timeout 200 python tools/bench_train.py --batch $B --steps 20 --warmup 4 > $O/train_b${B}_v2.json 2> $O/train_b${B}_v2.err; cat $O/train_b${B}_v2.json
done
FSR_TRAIN_OVERLAP=0 timeout 200 python tools/bench_train.py --batch 64 --steps 20 --warmup 4 > $O/train_b64_v2_nooverlap.json 2>/dev/null; cat $O/train_b64_v2_nooverlap.json
FSR_IN_BWD_FUSED=0 timeout 200 python tools/bench_train.py --batch 64 --steps 20 --warmup 4 > $O/train_b64_v2_inbwd2pass.json 2>/dev/null; cat $O/train_b64_v2_inbwd2pass.json
FSR_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_trainstep_b64_v2_eager3.csv \
    python tools/bench_train.py --batch 64 --steps 1 --warmup 2 > $O/ncu_train_v2.log 2>&1
FSR_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_gen_ws_kernel -s 60 -c 12 -o $O/genws_v2_full -f \
    python tools/bench_train.py --batch 64 --steps 1 --warmup 1 > $O/ncu_genws_v2.log 2>&1
ls -la $O | tail -8
