mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vae" > gpurun_out/r2_t_vae.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_vae.log
grep -E "passed|failed|Error|assert" gpurun_out/r2_t_vae.log | head -10
timeout 300 python tools/vae_bench.py --batch 32 > gpurun_out/r2_vae_bench.log 2>&1; cat gpurun_out/r2_vae_bench.log | grep -v "^Working\|^making"
