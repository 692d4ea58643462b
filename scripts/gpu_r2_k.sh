#!/bin/bash
# round 2, call K: with the pair kernels' operand prefetch off by default -- which prefetch hurt, and re-tune around it
mkdir -p gpurun_out
run() { # tag, args...
  tag=$1; shift
  timeout 400 python bench.py --no-extras --no-cpu-baseline --latency-runs 20 "$@" > gpurun_out/r2k_bench_$tag.json 2> gpurun_out/r2k_bench_$tag.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2k_bench_$tag.json"))
    print("$tag rc=$rc value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2))
except Exception as e:
    print("$tag rc=$rc parse failed", e)
PY
}
run default
run fwd1 --tune pair_prefetch_fwd=1
run bwd1 --tune pair_prefetch_bwd=1
run both1 --tune pair_prefetch_fwd=1 --tune pair_prefetch_bwd=1
run acc0 --tune acc_prefetch=0
run k8 --tune pair_k1=8 --tune pair_k2=8 --tune pair_k3=8
run k32 --tune pair_k1=32 --tune pair_k2=32 --tune pair_k3=32
run lat_h3w3 --tune prover_rounds_h=3 --tune prover_rounds_w=3
run lat_h0w3 --tune prover_rounds_h=0 --tune prover_rounds_w=3
run lat_h2w2 --tune prover_rounds_h=2 --tune prover_rounds_w=2
run lat_h1w1 --tune prover_rounds_h=1 --tune prover_rounds_w=1
run p6 --provers 6
run p8b16 --provers 8 --batch 16
run default_again
