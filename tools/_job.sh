mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "vae or sample_to_dir" > gpurun_out/r2_t_vae.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_vae.log
tail -15 gpurun_out/r2_t_vae.log | cut -c1-250
timeout 300 python tools/vae_bench.py --batch 32 --iters 3 > gpurun_out/r2_vae_bench4.log 2>&1; echo "rc=$?" >> gpurun_out/r2_vae_bench4.log
grep -E "libpdm|rc=" gpurun_out/r2_vae_bench4.log | cut -c1-300
