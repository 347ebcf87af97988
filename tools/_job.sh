mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t_all9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_all9.log
tail -8 gpurun_out/r2_t_all9.log | cut -c1-250
grep -q "rc=0" gpurun_out/r2_t_all9.log || exit 0
for v in 0 1 0 1; do
if [ $v = 1 ]; then export PDM_NO_FC2_ZC_FUSION=1; else unset PDM_NO_FC2_ZC_FUSION; fi
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extra-configs --no-kernel-profile 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nofusion=$v', d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'])"
done
