set -x
mkdir -p gpurun_out /tmp/prof
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-extra-configs --no-kernel-profile"
S="$B --nfe 10"
$S > gpurun_out/r02_plain_nfe10.log 2>&1 &&
ncu --set full --clock-control none -k regex:'embed_extras|embed_patch|rowstats_convert|copy_rows|head_token|conv3x3|update_kernel' -s 22 -c 11 -o /tmp/prof/r02_prof_hbm $S > gpurun_out/r02_ncu_hbm.log 2>&1
ncu --set full --clock-control none -k regex:ln_rstd -s 60 -c 1 -o /tmp/prof/r02_prof_lnrstd $S > gpurun_out/r02_ncu_lnrstd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 260 -c 16 -o /tmp/prof/r02_prof_gemm_small $S > gpurun_out/r02_ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc3 -s 60 -c 2 -o /tmp/prof/r02_prof_attn_small $S > gpurun_out/r02_ncu_attn.log 2>&1
python tools/kernel_bench.py --only layernorm > gpurun_out/r02_plain_ln.log 2>&1 &&
ncu --set full --clock-control none -k regex:layernorm_kernel -s 2 -c 1 -o /tmp/prof/r02_prof_layernorm python tools/kernel_bench.py --only layernorm > gpurun_out/r02_ncu_ln.log 2>&1
L="$B --nfe 6 --config large"
$L > gpurun_out/r02_plain_large.log 2>&1 &&
ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 200 -c 12 -o /tmp/prof/r02_prof_gemm_large $L > gpurun_out/r02_ncu_gemm_large.log 2>&1
# summaries (small) -> gpurun_out; raw reports only where they are small
python tools/summarise_ncu.py /tmp/prof/r02_prof_hbm.ncu-rep gpurun_out/r02_hbm_kernels_ncu.md "r02 -- ncu --set full of the HBM-bound kernels inside bench.py (config 2, 10 NFE)" "ncu --set full --clock-control none -k regex:'embed_extras|embed_patch|rowstats_convert|copy_rows|head_token|conv3x3|update_kernel' -s 22 -c 11 $S"
python tools/summarise_ncu.py /tmp/prof/r02_prof_lnrstd.ncu-rep gpurun_out/r02_lnrstd_ncu.md "r02 -- ln_rstd_kernel" "ncu --set full -k regex:ln_rstd -s 60 -c 1 $S"
python tools/summarise_ncu.py /tmp/prof/r02_prof_layernorm.ncu-rep gpurun_out/r02_layernorm_ncu.md "r02 -- layernorm_kernel (fp32 -> bf16, tools/kernel_bench.py --only layernorm)" "ncu --set full -k regex:layernorm_kernel -s 2 -c 1 python tools/kernel_bench.py --only layernorm"
python tools/summarise_ncu.py /tmp/prof/r02_prof_gemm_small.ncu-rep gpurun_out/r02_gemm_small_ncu.md "r02 -- gemm_tc_kernel inside bench.py (config 2)" "ncu --set full --import-source on -k regex:gemm_tc_kernel -s 260 -c 16 $S" gpurun_out/r02_gemm_traffic_small.json gemm_tc_kernel
python tools/summarise_ncu.py /tmp/prof/r02_prof_gemm_large.ncu-rep gpurun_out/r02_gemm_large_ncu.md "r02 -- gemm_tc_kernel inside bench.py --config large" "ncu --set full -k regex:gemm_tc_kernel -s 200 -c 12 $L" gpurun_out/r02_gemm_traffic_large.json gemm_tc_kernel
python tools/summarise_ncu.py /tmp/prof/r02_prof_attn_small.ncu-rep gpurun_out/r02_attention_ncu.md "r02 -- attention_tc3_kernel inside bench.py (config 2)" "ncu --set full --import-source on -k regex:attention_tc3 -s 60 -c 2 $S"
python tools/ncu_stall_table.py /tmp/prof/r02_prof_gemm_small.ncu-rep gpurun_out/r02_gemm_small_stalls.md "r02 -- warp-stall breakdown, gemm_tc_kernel launches inside bench.py (config 2)" 16
python tools/ncu_stall_table.py /tmp/prof/r02_prof_attn_small.ncu-rep gpurun_out/r02_attention_stalls.md "r02 -- warp-stall breakdown, attention_tc3_kernel inside bench.py (config 2)" 2
for f in r02_prof_hbm r02_prof_lnrstd r02_prof_layernorm r02_prof_attn_small; do sz=$(stat -c %s /tmp/prof/$f.ncu-rep); if [ $sz -lt 12000000 ]; then cp /tmp/prof/$f.ncu-rep gpurun_out/; fi; done
ls -la /tmp/prof gpurun_out/r02_*
# launch list of the default bench command (one full 50-NFE step)
$B > gpurun_out/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 46400 -c 11700 --csv --log-file gpurun_out/r02_bench_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
tail -2 gpurun_out/r02_ncu_launches.log; ls -la gpurun_out/r02_bench_launches.csv
timeout 600 python bench.py --method multistep --no-extra-configs --no-cpu-baseline --no-eager-baseline --steps 3 --warmup 3 > gpurun_out/r2_bench_multistep.json 2> gpurun_out/r2_bench_multistep.err
cat gpurun_out/r2_bench_multistep.json | cut -c1-400
du -sh gpurun_out
