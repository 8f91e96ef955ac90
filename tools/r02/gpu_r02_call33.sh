#!/bin/bash
# launch list of the pixel-pair path (n_filters = 32, L = 8, batch 32, 180x320): second forward only
O=gpurun_out/r02
mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_pairs_f32.csv python tools/profile_pairs.py 2 > $O/ncu_pairs.log 2>&1
python tools/agg_launches.py $O/launches_pairs_f32.csv | head -20
