#!/bin/bash
# round 2, GPU call 6 (2 GPUs): data-parallel GAN step through fsr_nccl_* inside the step graph; bench at N=2.
O=gpurun_out/r02
mkdir -p $O
nvidia-smi -L > $O/smi_2gpu.txt
timeout 600 python -m pytest tests/test_ddp_gpu.py -q -x -s > $O/t_ddp.log 2>&1; echo "rc=$?" >> $O/t_ddp.log; tail -25 $O/t_ddp.log
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench rc=$?"; tail -c 3000 $O/bench_n2.json; tail -5 $O/bench_n2.err
