mkdir -p gpurun_out /tmp/prof
S="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-extra-configs --no-kernel-profile --nfe 10"
timeout 300 $S > /tmp/prof/plain.json 2>/tmp/prof/plain.err; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'im2col_patch|embed_extras_emit|conv3x3_tok|update_kernel|gemm_tc_kernelILi2ELi7E|gemm_tc_kernel<2, 7>' -s 14 -c 7 -o /tmp/prof/r02_prof_hbm2 $S > /tmp/prof/ncu.log 2>&1; echo "ncu rc=$?"
ls -la /tmp/prof/
python tools/summarise_ncu.py /tmp/prof/r02_prof_hbm2.ncu-rep gpurun_out/r02_hbm_kernels_v2_ncu.md "r02 -- embed / head / update kernels of the end-of-round bf16 path (ncu --set full inside bench.py, config 2)" "ncu --set full --clock-control none -k regex:'im2col_patch|embed_extras_emit|conv3x3_tok|update_kernel|gemm_tc_kernel<2, 7>' -s 14 -c 7 $S" 2>&1 | tail -3
cat gpurun_out/r02_hbm_kernels_v2_ncu.md | cut -c1-400
