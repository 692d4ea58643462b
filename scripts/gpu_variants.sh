# throughput of library build variants / prover counts:  VARIANTS="name:provers ..." (name "" = default build)
for V in ${VARIANTS}; do
  NAME=${V%%:*}; P=${V##*:}
  if [ -n "$NAME" ]; then export NZCP_LIB_PATH=$PWD/nzcp_circom_b200/libnzcp_prover_$NAME.so; else unset NZCP_LIB_PATH; fi
  python bench.py --steps 3 --warmup 3 --batch 8 --provers $P --no-cpu-baseline > gpurun_out/var_${NAME}_p$P.json 2> gpurun_out/var_${NAME}_p$P.err
  python - <<PY
import json
d=json.load(open('gpurun_out/var_${NAME}_p$P.json'))
print('variant=[$NAME] P=$P value %.1f e2e %.1f p50 %.2f' % (d['value'], d['e2e']['value'], d['p50_latency_ms']), 'acc', {k: round(v,2) for k,v in d['msm']['accumulate_ms'].items()})
PY
done
