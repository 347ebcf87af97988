mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge.py -m gpu -q > gpurun_out/r2_t_edge.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_edge.log
tail -25 gpurun_out/r2_t_edge.log | cut -c1-220
