#!/bin/bash
# round 2, call E: pair-round combinations (witness vs H), then the whole GPU suite on the final build
mkdir -p gpurun_out
run() { # tag, args...
  tag=$1; shift
  timeout 400 python bench.py --no-extras --no-cpu-baseline "$@" > gpurun_out/r2e_bench_$tag.json 2> gpurun_out/r2e_bench_$tag.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2e_bench_$tag.json"))
    print("$tag rc=$rc value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2), "acc", {k: round(v,2) for k,v in d["msm"]["accumulate_ms"].items()})
except Exception as e:
    print("$tag rc=$rc parse failed", e)
PY
}
run default
run w1 --tune prover_rounds_w=1
run w2 --tune prover_rounds_w=2
run w3 --tune prover_rounds_w=3
run h1w2 --tune prover_rounds_h=1 --tune prover_rounds_w=2
run h3w3 --tune prover_rounds_h=3 --tune prover_rounds_w=3
run w2k32 --tune prover_rounds_w=2 --tune pair_k1=32 --tune pair_k2=32
run w2p6 --tune prover_rounds_w=2 --provers 6
run default_p6 --provers 6
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -6 gpurun_out/r2e_pytest_gpu.log
