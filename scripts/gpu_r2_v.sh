#!/bin/bash
# round 2, call V: the G2 MSM with its own rounds-free (or shorter) sort vs sharing the witness sort's three rounds
mkdir -p gpurun_out
run() { # tag, args...
  tag=$1; shift
  timeout 400 python bench.py --no-extras --latency-runs 30 "$@" > gpurun_out/r2v_bench_$tag.json 2> gpurun_out/r2v_bench_$tag.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2v_bench_$tag.json"))
    print("$tag rc=$rc value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2), "min", round(d["latency_ms"]["min"],2), "verifies", d.get("proof_verifies"), "matches_cpu_port", d.get("proof_matches_cpu_port"), "acc_b2", round(d["msm"]["accumulate_ms"]["b2"],2))
except Exception as e:
    print("$tag rc=$rc parse failed", e)
PY
}
run default
run b2_0 --tune prover_rounds_b2=0
run b2_1 --tune prover_rounds_b2=1
run default_b
run b2_0_b --tune prover_rounds_b2=0
