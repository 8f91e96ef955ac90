#!/bin/bash
# round 2, GPU call 2: fused residual chain + CTA-pair up conv in the product; single-graph train step.
O=gpurun_out/r02
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "not float32" --durations=8 > $O/t_all2.log 2>&1; echo "rc=$?" >> $O/t_all2.log
tail -4 $O/t_all2.log
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
head -c 400 $O/bench_n1.json; echo
FSR_UP_2CTA=0 FSR_FUSE_RES=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline > $O/bench_n1_r01path.json 2> $O/bench_n1_r01path.err
FSR_UP_2CTA=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline > $O/bench_n1_no2cta.json 2> $O/bench_n1_no2cta.err
FSR_FUSE_RES=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline > $O/bench_n1_nofuseres.json 2> $O/bench_n1_nofuseres.err
FSR_TRAIN_OVERLAP=0 timeout 200 python tools/bench_train.py --batch 64 --steps 10 --warmup 4 > $O/train_b64_nooverlap.json 2> $O/train_b64_nooverlap.err
timeout 200 python tools/bench_train.py --batch 64 --steps 10 --warmup 4 > $O/train_b64.json 2> $O/train_b64.err
cat $O/train_b64.json $O/train_b64_nooverlap.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 22 -c 24 --csv --log-file $O/launches_generator_b32_180x320.csv \
    python tools/profile_step.py 2 > $O/ncu_gen.log 2>&1
FSR_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_trainstep_b64_eager3.csv \
    python tools/bench_train.py --batch 64 --steps 1 --warmup 2 > $O/ncu_train.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:up_2cta -s 2 -c 2 -o $O/up2cta_full -f \
    python tools/profile_step.py 2 > $O/ncu_up.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_c64_kernel -s 17 -c 3 -o $O/resconv_full -f \
    python tools/profile_step.py 2 > $O/ncu_res.log 2>&1
ls -la $O
