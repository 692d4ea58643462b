#!/bin/bash
# round 2, call M (1 GPU): final build -- whole GPU suite, bench + reference arm of record, launch list + pipe time of a proof
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/r2m_pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_ref.json 2> gpurun_out/r2m_ref.err; echo "ref rc=$?"
( time python bench.py ) > gpurun_out/r2m_bench_default.json 2> gpurun_out/r2m_bench_default.err; echo "bench rc=$?"; tail -4 gpurun_out/r2m_bench_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2m_bench_default.json"))
    print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2), round(d["p50_latency_ms_pageable"],2), d["latency_ms"], "launches", d["gpu_launches"], d.get("proof_verifies"), d.get("proof_matches_cpu_port"), "frac", round(d["roofline"]["frac"],3), round(d["roofline_step"]["frac"],3), round(d["roofline_step"]["frac_without_reduction_term"],3))
    r=json.load(open("gpurun_out/r2m_ref.json")); print("ref", r["value"], r["cpu_baseline"]["cores"], r["config"]==d["config"])
    print(json.dumps(d["extras"])[:1200])
except Exception as e:
    print("bench parse failed", e)
PY
python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras --latency-runs 2 > gpurun_out/r2m_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r2m_launches_pipe.csv python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras --latency-runs 2 > gpurun_out/r2m_ncu.log 2>&1; echo "ncu list rc=$?"
