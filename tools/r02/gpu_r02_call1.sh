mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02/smi.txt 2>&1
FSR_REPORT_DIR=gpurun_out/r02 timeout 600 python -m pytest tests/test_baseline_configs_gpu.py -q -s -k "not float32" > gpurun_out/r02/t_base.log 2>&1
echo "base rc=$?" >> gpurun_out/r02/t_base.log
FSR_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -q -s > gpurun_out/r02/t_exp.log 2>&1
echo "exp rc=$?" >> gpurun_out/r02/t_exp.log
timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_baseline_configs_gpu.py > gpurun_out/r02/t_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02/t_all.log
tail -5 gpurun_out/r02/t_base.log gpurun_out/r02/t_exp.log gpurun_out/r02/t_all.log
