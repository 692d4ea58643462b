# throughput vs proofs per step and provers in flight: CASES="batch:provers ..."
for V in ${CASES:-12:4 24:4 48:4 24:6 48:6 48:8}; do
  B=${V%%:*}; P=${V##*:}
  python bench.py --steps 3 --warmup 3 --batch $B --provers $P --no-cpu-baseline > gpurun_out/sweep_b${B}_p$P.json 2> gpurun_out/sweep_b${B}_p$P.err
  python - <<PY
import json
d=json.load(open('gpurun_out/sweep_b${B}_p$P.json'))
print('B=$B P=$P value %.1f e2e %.1f p50 %.2f ms/step %.1f' % (d['value'], d['e2e']['value'], d['p50_latency_ms'], d['ms_per_step']))
PY
done
