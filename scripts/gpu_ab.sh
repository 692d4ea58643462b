# A/B of an environment switch of the library: ENVNAME=... VALUES="0 1 0 1" (throughput, e2e, p50, lone-proof stage times)
for V in ${VALUES:-0 1 0 1}; do
env ${ENVNAME}=$V python bench.py --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_ab$V.json 2> gpurun_out/bench_ab$V.err; echo "${ENVNAME}=$V bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab$V.json'))
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'p50',round(d['p50_latency_ms'],2), 'total_ms', round(d['msm']['total_ms'],2), 'r1cs', round(d['stage_ms']['r1cs_eval'],3), 'ntt', round(d['stage_ms']['ntt_join'],2), 'msm_h', round(d['stage_ms']['msm_h'],2), 'sort', {k: round(v,2) for k,v in d['msm']['sort_ms'].items()}, 'acc', {k: round(v,2) for k,v in d['msm']['accumulate_ms'].items()})
PY
done
