#!/bin/bash
# ncu --set full of one fused InstanceNorm-backward launch (generator chain: 24x24, 64 channels, batch 64) and one instnorm_apply
O=gpurun_out/r02
mkdir -p $O
FSR_GRAPH=0 FSR_TRAIN_OVERLAP=0 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:instnorm_bwd_fused --launch-skip 45 --launch-count 1 \
   -o $O/in_bwd_fused_full -f python tools/bench_train.py --batch 64 --steps 1 --warmup 1 > $O/ncu_in_bwd.log 2>&1
tail -2 $O/ncu_in_bwd.log
python tools/ncu_summary.py $O/in_bwd_fused_full.ncu-rep | head -8
ls -la $O/in_bwd_fused_full.ncu-rep
