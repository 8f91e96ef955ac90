#!/bin/bash
# round 2, GPU call 8: mbarrier-poll backoff A/B on the 64->64 conv; reproducibility test; generator launch list.
O=gpurun_out/r02
mkdir -p $O
timeout 300 python -m pytest tests/test_train_step_gpu.py -q -x -k "reproducible or shard" > $O/t_call8.log 2>&1; echo "rc=$?" >> $O/t_call8.log; tail -3 $O/t_call8.log
for ns in 0 20 64 200; do
  FSR_BACKOFF_NS=$ns timeout 200 python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline > $O/bench_backoff_$ns.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("$O/bench_backoff_$ns.json")); r=d["roofline_resblock_conv"]
print("backoff $ns ns:", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms; res conv avg", round(r["avg_launch_ms"]*1e3,1), "us frac", round(r["frac"],3), "up", round(d["roofline"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"])
PY
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 53 -c 30 --csv --log-file $O/launches_generator_b32_180x320_v5.csv \
    python tools/profile_step.py 2 > $O/ncu_gen_v5.log 2>&1
