mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_h.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'clocks',d['clocks'])
print('roofline',d['roofline']['frac'],d['roofline']['achieved'],'attn',d['roofline_attention']['achieved'], d['roofline_attention']['share_of_step'])
print('ms_per_nnet_step',d['ms_per_nnet_step'],'frac',d['frac_of_bf16_peak'],'tf',d['model_tflops_per_gpu'])
print('vae',d.get('vae_decode'))
for k,v in d.get('configs',{}).items(): print(k, v.get('samples_per_s'), v.get('ms_per_nnet_step'), v.get('model_tflops_per_gpu'), v.get('frac_of_bf16_peak'), v.get('roofline',{}).get('achieved'), v.get('roofline',{}).get('frac'), v.get('roofline_attention',{}).get('achieved'), v.get('roofline_attention',{}).get('share_of_step'))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 38000 -c 7900 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-extra-configs --no-kernel-profile > gpurun_out/r02e_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02e_launches.csv
