#!/bin/bash
# round 2, GPU call 12 (2 GPUs): data-parallel check through fsr_nccl_* (graph) and through the torch.distributed fallback (eager).
O=gpurun_out/r02
mkdir -p $O
timeout 420 python -m pytest tests/test_ddp_gpu.py -q -x -s > $O/t_ddp2.log 2>&1; echo "rc=$?" >> $O/t_ddp2.log; grep -E "DDP CHECK|spread|passed|failed|rc=" $O/t_ddp2.log
