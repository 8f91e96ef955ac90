#!/bin/bash
# round 2, GPU call 9: sub-batch overlap (HBM-bound head / normalise passes under the tensor-bound CTA-pair up conv).
O=gpurun_out/r02
mkdir -p $O
for st in 1 2 3 4; do
  timeout 200 python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline --streams $st > $O/bench_streams_$st.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("$O/bench_streams_$st.json")); r=d["roofline_resblock_conv"]
print("streams $st:", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1), "res", round(r["avg_launch_ms"]*1e3,1), "up", round(d["roofline"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
