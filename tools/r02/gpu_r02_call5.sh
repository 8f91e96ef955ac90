#!/bin/bash
# round 2, GPU call 5: CTA-pair general conv, parallel wgrad reduce, 16-channel fused IN backward, parity-out IN apply, autograd bridge.
O=gpurun_out/r02
mkdir -p $O
timeout 300 python -m pytest tests/test_gen_2cta_gpu.py -q -x > $O/t_call5_2cta.log 2>&1; echo "rc=$?" >> $O/t_call5_2cta.log; tail -3 $O/t_call5_2cta.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gen_2cta_gpu.py --durations=5 > $O/t_call5.log 2>&1; echo "rc=$?" >> $O/t_call5.log
tail -12 $O/t_call5.log
for B in 64 32; do
  timeout 200 python tools/bench_train.py --batch $B --steps 20 --warmup 4 > $O/train_b${B}_v3.json 2> $O/train_b${B}_v3.err; cat $O/train_b${B}_v3.json
done
FSR_GEN_2CTA=0 timeout 200 python tools/bench_train.py --batch 64 --steps 20 --warmup 4 > $O/train_b64_v3_no2cta.json 2>/dev/null; cat $O/train_b64_v3_no2cta.json
FSR_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_trainstep_b64_v3_eager3.csv \
    python tools/bench_train.py --batch 64 --steps 1 --warmup 2 > $O/ncu_train_v3.log 2>&1
FSR_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_gen_2cta -s 40 -c 10 -o $O/gen2cta_full -f \
    python tools/bench_train.py --batch 64 --steps 1 --warmup 1 > $O/ncu_gen2cta.log 2>&1
ls -la $O | tail -5
