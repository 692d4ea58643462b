# A/B: stream priority of the H chain on/off (throughput, e2e, p50) after the parity suite
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
for V in 0 1 0 1; do
NZCP_NO_STREAM_PRIORITY=$V python bench.py --no-cpu-baseline > gpurun_out/bench_prio$V.json 2> gpurun_out/bench_prio$V.err; echo "noprio=$V bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_prio$V.json'))
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'p50',round(d['p50_latency_ms'],2), 'total_ms', round(d['msm']['total_ms'],2), 'r1cs', round(d['stage_ms']['r1cs_eval'],3), 'ntt', round(d['stage_ms']['ntt_join'],2), 'msm_h', round(d['stage_ms']['msm_h'],2), 'sort', d['msm']['sort_ms'])
PY
done
