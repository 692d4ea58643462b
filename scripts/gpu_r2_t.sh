#!/bin/bash
# round 2, call T: thinner blocks for the shared-memory histogram passes of the bucket sort (placement beside other kernels)
mkdir -p gpurun_out
run() { # tag, args...
  tag=$1; shift
  timeout 400 python bench.py --no-extras --no-cpu-baseline --latency-runs 30 "$@" > gpurun_out/r2t_bench_$tag.json 2> gpurun_out/r2t_bench_$tag.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2t_bench_$tag.json"))
    print("$tag rc=$rc value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2), "min", round(d["latency_ms"]["min"],2), "sort", {k: round(v,2) for k,v in d["msm"]["sort_ms"].items()}, "total", round(d["msm"]["total_ms"],2), d.get("proof_verifies"))
except Exception as e:
    print("$tag rc=$rc parse failed", e)
PY
}
run t1024
run t512 --tune sort_threads=512
run t256 --tune sort_threads=256
run t128 --tune sort_threads=128
run t1024_b
run t256_b --tune sort_threads=256
