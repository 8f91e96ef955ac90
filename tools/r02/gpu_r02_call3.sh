#!/bin/bash
# round 2, GPU call 3: 2-CTA up conv with the 8-warp TMA-store epilogue; precise (float32) generator mode.
O=gpurun_out/r02
mkdir -p $O
timeout 300 python -m pytest tests/test_fused_chain_gpu.py tests/test_baseline_configs_gpu.py tests/test_generator_gpu.py -q -x -s -k "not train and not pretrain" > $O/t_call3.log 2>&1; echo "rc=$?" >> $O/t_call3.log
grep -E "max-abs|passed|failed|rc=" $O/t_call3.log | tail -20
timeout 300 python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline > $O/bench_n1_v2.json 2> $O/bench_n1_v2.err
FSR_UP_2CTA=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline > $O/bench_n1_v2_no2cta.json 2> $O/bench_n1_v2_no2cta.err
python - <<'PY'
import json
for f in ("bench_n1_v2","bench_n1_v2_no2cta"):
    try:
        d=json.load(open(f"gpurun_out/r02/{f}.json")); print(f, round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms  up1", round(d["roofline"]["avg_launch_ms"],3), "ms frac", round(d["roofline"]["frac"],3), "res", round(d["roofline_resblock_conv"]["avg_launch_ms"],4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e: print(f, "ERR", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:up_2cta -s 2 -c 2 -o $O/up2cta_v2_full -f \
    python tools/profile_step.py 2 > $O/ncu_up_v2.log 2>&1
