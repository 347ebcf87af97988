mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n2.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n2.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['cpu_baseline'], d['torch_eager_b200'])
for k,v in d['configs'].items(): print(k, v['samples_per_s'], v['global_batch'], v['scaling'], v['config']['batch_per_gpu'])
"
