# GPU parity suite + a short bench at 1..3 provers in flight (summary to stdout, JSON lines to gpurun_out/)
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_gpu.log
python -c "
from nzcp_circom_b200 import api
m = api.intpipe_modes(0, 4096)
for k, v in m.items(): print('%-16s %8.3f T/s' % (k, v / 1e12))
" 2>&1 | tee gpurun_out/intpipe_modes.txt
for P in ${PROVERS:-1 3}; do python bench.py --steps 3 --warmup 3 --batch 6 --provers $P --no-cpu-baseline > gpurun_out/bench_p$P.json 2> gpurun_out/bench_p$P.err; echo rc=$?; python - <<PY
import json
d=json.load(open('gpurun_out/bench_p$P.json'))
print('P=$P', 'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'p50',round(d['p50_latency_ms'],2), 'launches', d['gpu_launches'])
print(d['stage_ms']); print(d['msm']); print(d['roofline']['frac'], d['roofline']['launch_ms'])
PY
done
