#!/bin/bash
# 8 GPUs: the driver's SCALE launch at N=8 with the side-stream schedule, short.
O=gpurun_out/r02
mkdir -p $O
NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 \
    bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n8_v2.json 2> $O/bench_n8_v2.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_n8_v2.json").read().strip().splitlines()[-1])
print("N=8 value", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1))
print(json.dumps(d["train_step"], indent=1)[:1400])
PY
tail -3 $O/bench_n8_v2.err
