mkdir -p gpurun_out
for lib in libpdm.so libpdm_p0x0000.so libpdm_p0x0808.so libpdm_p0x8080.so libpdm_p0xaaaa.so; do
echo "== $lib"
PDM_LIB=$PWD/panopticdiffusionmodels_b200/$lib python tools/kernel_bench.py --only attention --attn-nb 512
PDM_LIB=$PWD/panopticdiffusionmodels_b200/$lib python tools/kernel_bench.py --only attention --attn-nb 512 --attn-L 334
done > gpurun_out/r2_attn_poly.log 2>&1
cat gpurun_out/r2_attn_poly.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t_all2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_all2.log
tail -15 gpurun_out/r2_t_all2.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_a.err
cat gpurun_out/r2_bench_a.json
