mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "rc=$?"
tail -3 gpurun_out/r2_bench_n4.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n4.json').read().strip().splitlines()[-1])
print('value',d['value'],'n',d['n_gpus'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'clocks',d['clocks']['sm_mhz'])
for k,v in d.get('configs',{}).items(): print(k, v.get('samples_per_s'), v.get('global_batch'), v.get('scaling'))
PY
