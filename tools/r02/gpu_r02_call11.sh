#!/bin/bash
# round 2, GPU call 11: one-launch weight packs, parity-layout dy in the InstanceNorm backward; full suite + train bench.
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/t_call11.log 2>&1; echo "rc=$?" >> $O/t_call11.log
tail -5 $O/t_call11.log
for B in 64 32; do
  timeout 200 python tools/bench_train.py --batch $B --steps 20 --warmup 4 > $O/train_b${B}_v5.json 2> $O/train_b${B}_v5.err; cat $O/train_b${B}_v5.json
done
FSR_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_trainstep_b64_v5_eager3.csv \
    python tools/bench_train.py --batch 64 --steps 1 --warmup 2 > $O/ncu_train_v5.log 2>&1
