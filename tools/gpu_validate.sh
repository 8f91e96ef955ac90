#!/bin/bash
# One gpurun call that validates a change end to end - what the driver runs at round end, plus the launch lists:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh'
# parity suite -> smoke -> headline bench (ours, reference arm) -> train bench -> ncu launch lists.
# (tools/r02/ keeps the exact scripts of every GPU call of round 2.)
O=gpurun_out/validate
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log; tail -4 $O/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python bench.py --steps 30 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; head -c 400 $O/bench_n1.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; head -c 300 $O/bench_reference.json; echo
python tools/bench_train.py --batch 64 --steps 20 --warmup 4 > $O/train_b64.json 2> $O/train_b64.err; cat $O/train_b64.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 53 -c 30 --csv --log-file $O/launches_generator_b32_180x320.csv \
    python tools/profile_step.py 2 > $O/ncu_gen.log 2>&1
FSR_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_trainstep_b64_eager3.csv \
    python tools/bench_train.py --batch 64 --steps 1 --warmup 2 > $O/ncu_train.log 2>&1
