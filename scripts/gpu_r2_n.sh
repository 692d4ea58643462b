#!/bin/bash
# round 2, call N (1 GPU): bench + reference arm of record on the final build
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_ref.json 2> gpurun_out/r2n_ref.err; echo "ref rc=$?"
( time python bench.py ) > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err; echo "bench rc=$?"; tail -4 gpurun_out/r2n_bench_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2n_bench_default.json"))
    print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2), round(d["p50_latency_ms_pageable"],2), "launches", d["gpu_launches"], d.get("proof_verifies"), d.get("proof_matches_cpu_port"))
    r=d["roofline"]; print("roofline frac", round(r["frac"],3), "launch_ms", r["launch_ms"], "in_proof", r["in_proof_ms"], "xyzz", r.get("xyzz_kernel"))
    print("step", round(d["roofline_step"]["frac"],3), round(d["roofline_step"]["frac_without_reduction_term"],3))
    rr=json.load(open("gpurun_out/r2n_ref.json")); print("ref", rr["value"], rr["cpu_baseline"]["cores"], rr["config"]==d["config"])
except Exception as e:
    print("bench parse failed", e)
PY
