#!/bin/bash
# round 2, call J: pair-round kernel variants (operand prefetch target, forward-kernel occupancy) as separate builds
mkdir -p gpurun_out
run() { # tag, lib, args...
  tag=$1; lib=$2; shift; shift
  NZCP_LIB_PATH=$lib timeout 400 python bench.py --no-extras --no-cpu-baseline "$@" > gpurun_out/r2j_bench_$tag.json 2> gpurun_out/r2j_bench_$tag.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2j_bench_$tag.json"))
    print("$tag rc=$rc value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2))
except Exception as e:
    print("$tag rc=$rc parse failed", e)
PY
}
L=$PWD/nzcp_circom_b200
for rep in 1 2; do
run default_$rep $L/libnzcp_prover.so
for v in pf0 pf2 occ8; do
  if [ -f $L/libnzcp_prover_$v.so ]; then run ${v}_$rep $L/libnzcp_prover_$v.so; fi
done
done
