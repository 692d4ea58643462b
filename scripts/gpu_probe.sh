# GPU parity suite, default bench, and the pipe-overlap probes (summary to stdout, files to gpurun_out/)
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
python -c "
from nzcp_circom_b200 import api
for k, v in api.pipe_probe(0, 4096).items(): print('%-22s %8.3f T/s' % (k, v / 1e12))
for k, v in api.intpipe_modes(0, 4096).items(): print('%-22s %8.3f T/s' % (k, v / 1e12))
" 2>&1 | tee gpurun_out/pipe_probe.txt
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'p50',round(d['p50_latency_ms'],2), 'launches', d['gpu_launches'])
print(d['stage_ms']); print(d['msm']); print(d['roofline']['frac'], d['roofline']['launch_ms']); print(d['roofline_ntt'].get('alone')); print(d.get('proof_matches_cpu_port'))
PY
