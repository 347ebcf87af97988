mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t_all4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_all4.log
tail -5 gpurun_out/r2_t_all4.log | cut -c1-250
timeout 600 python bench.py --no-extra-configs --no-cpu-baseline --no-eager-baseline --steps 3 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_c.json')); print(d['value'], d['ms_per_step'])
print({k:(v['avg_ms'],v['share']) for k,v in d['kernels'].items()})
"
S="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-extra-configs --no-kernel-profile --nfe 4"
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 260 --csv --log-file gpurun_out/r02c_launches.csv $S > gpurun_out/r02c_ncu_launches.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r02c_launches.csv') if not l.startswith("==")) if r]
h=rows[0]; ik,iv=h.index("Kernel Name"),h.index("Metric Value")
seen=0
for r in rows[1:]:
    n=r[ik].split("(")[0][-45:]
    if ("conv3x3" in n or "im2col" in n or "embed_extras" in n or "update" in n or "gemm_tc_kernel<2, 0>" in n) and seen<14:
        print(n, r[iv]); seen+=1
PY
