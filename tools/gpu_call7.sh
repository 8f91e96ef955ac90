#!/bin/bash
mkdir -p gpurun_out
timeout 70 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/t_all.log
tail -1 gpurun_out/t_all.log
timeout 100 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
head -c 230 gpurun_out/bench_n1.json; echo
