#!/bin/bash
# round 2, call I (1 GPU): whole GPU suite on the final build, bench + reference arm of record, launch list of a throughput-mode proof
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/r2i_pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_ref.json 2> gpurun_out/r2i_ref.err; echo "ref rc=$?"
( time python bench.py ) > gpurun_out/r2i_bench_default.json 2> gpurun_out/r2i_bench_default.err; echo "bench rc=$?"; tail -4 gpurun_out/r2i_bench_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2i_bench_default.json"))
    print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2), round(d["p50_latency_ms_pageable"],2), "launches", d["gpu_launches"], d.get("proof_verifies"), d.get("proof_matches_cpu_port"), "frac", round(d["roofline"]["frac"],3), round(d["roofline_step"]["frac"],3), round(d["roofline_step"]["frac_without_reduction_term"],3))
    r=json.load(open("gpurun_out/r2i_ref.json")); print("ref", r["value"], r["cpu_baseline"]["cores"], r["config"]==d["config"])
except Exception as e:
    print("bench parse failed", e)
PY
python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras > gpurun_out/r2i_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 1100 --csv --log-file gpurun_out/r2i_launches_pipe_throughput.csv python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras > gpurun_out/r2i_ncu.log 2>&1; echo "ncu list rc=$?"
