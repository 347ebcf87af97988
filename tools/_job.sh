mkdir -p gpurun_out
export PDM_ATTN_V4=1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -v -k "attention" > gpurun_out/r2_t_attn4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_attn4.log
tail -6 gpurun_out/r2_t_attn4.log | cut -c1-200
grep -q "rc=0" gpurun_out/r2_t_attn4.log || exit 0
timeout 200 python tools/attn_stress.py 3 > gpurun_out/r2_stress4.log 2>&1; grep -c BAD gpurun_out/r2_stress4.log; tail -2 gpurun_out/r2_stress4.log
for v in 1 0; do
if [ $v = 1 ]; then export PDM_ATTN_V4=1; else unset PDM_ATTN_V4; fi
echo "== V4=$v"
timeout 100 python tools/kernel_bench.py --only attention --attn-nb 512
timeout 100 python tools/kernel_bench.py --only attention --attn-nb 512 --attn-L 334
timeout 100 python tools/kernel_bench.py --only attention --attn-nb 64 --attn-L 2126
done > gpurun_out/r2_attn4.log 2>&1
grep -E "==|kernel" gpurun_out/r2_attn4.log
export PDM_ATTN_V4=1
PDM_LIB=$PWD/panopticdiffusionmodels_b200/libpdm_trace.so PDM_ATTN_TRACE_FILE=gpurun_out/r2_trace590d.bin timeout 120 python tools/kernel_bench.py --only attention --attn-nb 128 > gpurun_out/r2_trace590d.log 2>&1
