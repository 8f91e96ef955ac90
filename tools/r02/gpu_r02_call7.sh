#!/bin/bash
# round 2, GPU call 7: full suite, train bench after the parallel wgrad reduce, headline bench, launch lists, incumbent with torch.compile.
O=gpurun_out/r02
mkdir -p $O
timeout 300 python -m pytest tests/test_flat_conv_gpu.py -q > $O/t_call7_flat.log 2>&1; FLAT_RC=$?; echo "rc=$FLAT_RC" >> $O/t_call7_flat.log; tail -4 $O/t_call7_flat.log
if [ "$FLAT_RC" != "0" ]; then export FSR_VGG_FLAT=0; echo "flat conv tests failed: FSR_VGG_FLAT=0 for the rest of this call"; fi
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 --deselect tests/test_flat_conv_gpu.py > $O/t_call7.log 2>&1; echo "rc=$?" >> $O/t_call7.log
tail -6 $O/t_call7.log
FSR_VGG_FLAT=0 timeout 200 python tools/bench_train.py --batch 64 --steps 20 --warmup 4 > $O/train_b64_v4_noflat.json 2>/dev/null; cat $O/train_b64_v4_noflat.json
for B in 64 32; do
  timeout 200 python tools/bench_train.py --batch $B --steps 20 --warmup 4 > $O/train_b${B}_v4.json 2> $O/train_b${B}_v4.err; cat $O/train_b${B}_v4.json
done
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench_n1_v4.json 2> $O/bench_n1_v4.err; head -c 600 $O/bench_n1_v4.json; echo
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 32 --csv --log-file $O/launches_generator_b32_180x320_v4.csv \
    python tools/profile_step.py 2 > $O/ncu_gen_v4.log 2>&1
FSR_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_trainstep_b64_v4_eager3.csv \
    python tools/bench_train.py --batch 64 --steps 1 --warmup 2 > $O/ncu_train_v4.log 2>&1
timeout 420 python tools/incumbent.py --steps 10 --no-train > $O/incumbent_gen_compile.json 2> $O/incumbent_gen_compile.err; cat $O/incumbent_gen_compile.json
timeout 200 python tools/incumbent.py --steps 10 --no-compile > $O/incumbent_eager.json 2> $O/incumbent_eager.err; cat $O/incumbent_eager.json
