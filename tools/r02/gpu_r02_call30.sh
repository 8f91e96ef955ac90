#!/bin/bash
# 2 GPUs: does the exchange hide better with the D(sr) adversarial branch on the main stream (side = D step + all-reduce + update)?
O=gpurun_out/r02
mkdir -p $O
for a in 1 0 1 0; do
FSR_ADV_SIDE=$a NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2964$a \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n2_adv$a.json 2> $O/bench_n2_adv$a.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02/bench_n2_adv$a.json").read().strip().splitlines()[-1])
t=d["train_step"]
print("adv_side=$a step", round(t["ms_per_step"],3), "anchor", round(t["n1_anchor_b32_ms"],3), "eff", round(t["efficiency_vs_n1"],4), "diff", t["replica_max_diff"])
PY
done
