mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t_all6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_all6.log
tail -12 gpurun_out/r2_t_all6.log | cut -c1-250
timeout 600 python bench.py --no-extra-configs --no-cpu-baseline --no-eager-baseline --steps 3 --warmup 3 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_e.json')); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'])
print({k:(v['avg_ms'],v['share']) for k,v in d['kernels'].items()})
"
