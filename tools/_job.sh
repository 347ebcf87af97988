mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t_all7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_all7.log
tail -6 gpurun_out/r2_t_all7.log | cut -c1-250
timeout 900 python bench.py --no-eager-baseline --no-cpu-baseline > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_f.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_f.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'clocks',d['clocks'])
print('roofline',d['roofline']['frac'],d['roofline']['achieved'])
for k,v in d.get('configs',{}).items(): print(k, v.get('samples_per_s'), v.get('frac_of_bf16_peak'))
PY
