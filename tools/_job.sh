mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "vae or sample_to_dir or linear" > gpurun_out/r2_t_vae.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_vae.log
tail -15 gpurun_out/r2_t_vae.log | cut -c1-250
timeout 300 python tools/vae_bench.py --batch 32 --iters 3 > gpurun_out/r2_vae_bench5.log 2>&1; echo "rc=$?" >> gpurun_out/r2_vae_bench5.log
grep -E "libpdm|rc=" gpurun_out/r2_vae_bench5.log | cut -c1-300
PDM_GEMM_NO_N128=1 timeout 300 python tools/vae_bench.py --batch 32 --iters 3 2>&1 | grep '(libpdm)' | cut -c1-200
