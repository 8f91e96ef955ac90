#!/bin/bash
# 2 GPUs: data-parallel correctness test + the driver's SCALE launch at N=2 with the new stream schedule (D step / weight gradients on the side stream)
O=gpurun_out/r02
mkdir -p $O
timeout 400 python -m pytest tests/test_ddp_gpu.py -x -q -m gpu 2>&1 | tail -3
NCCL_DEBUG=WARN timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n2_v2.json 2> $O/bench_n2_v2.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_n2_v2.json").read().strip().splitlines()[-1])
print("N=2 value", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1))
print(json.dumps(d["train_step"], indent=1)[:1500])
PY
tail -3 $O/bench_n2_v2.err
