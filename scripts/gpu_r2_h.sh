#!/bin/bash
# round 2, call H: throughput-mode tuning -- additions per thread in the pair rounds, provers in flight, proofs per step
mkdir -p gpurun_out
run() { # tag, args...
  tag=$1; shift
  timeout 400 python bench.py --no-extras --no-cpu-baseline "$@" > gpurun_out/r2h_bench_$tag.json 2> gpurun_out/r2h_bench_$tag.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2h_bench_$tag.json"))
    print("$tag rc=$rc value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2))
except Exception as e:
    print("$tag rc=$rc parse failed", e)
PY
}
run default
run k8 --tune pair_k1=8 --tune pair_k2=8 --tune pair_k3=8
run k16_8_8 --tune pair_k1=16 --tune pair_k2=8 --tune pair_k3=8
run k32_16_16 --tune pair_k1=32 --tune pair_k2=16 --tune pair_k3=16
run k8_16_16 --tune pair_k1=8 --tune pair_k2=16 --tune pair_k3=16
run p3 --provers 3
run p5 --provers 5
run p6 --provers 6 --batch 12
run p8b16 --provers 8 --batch 16
run p6b24 --provers 6 --batch 24
run h2w3 --tune prover_rounds_h=2 --tune prover_rounds_w=3
run h0w3 --tune prover_rounds_h=0 --tune prover_rounds_w=3
