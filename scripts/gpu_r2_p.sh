#!/bin/bash
# round 2, call P (1 GPU): FINAL build -- whole GPU suite, smoke, reference arm, bench of record, launch list with pipe time + DRAM bytes
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/r2p_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2p_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_ref.json 2> gpurun_out/r2p_ref.err; echo "ref rc=$?"
( time python bench.py ) > gpurun_out/r2p_bench_default.json 2> gpurun_out/r2p_bench_default.err; echo "bench rc=$?"; tail -4 gpurun_out/r2p_bench_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2p_bench_default.json"))
    r=d["roofline"]
    print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2), round(d["p50_latency_ms_pageable"],2), d["latency_ms"], "launches", d["gpu_launches"], d.get("proof_verifies"), d.get("proof_matches_cpu_port"))
    print("roofline", round(r["frac"],3), r["launch_ms"], r["in_proof_ms"], (r.get("xyzz_kernel") or {}).get("launch_ms"), "step", round(d["roofline_step"]["frac"],3), round(d["roofline_step"]["frac_without_reduction_term"],3))
    rr=json.load(open("gpurun_out/r2p_ref.json")); print("ref", rr["value"], rr["cpu_baseline"]["cores"], rr["config"]==d["config"])
    print(json.dumps(d["extras"])[:1300])
except Exception as e:
    print("bench parse failed", e)
PY
python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras --latency-runs 2 > gpurun_out/r2p_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r2p_launches_pipe.csv python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras --latency-runs 2 > gpurun_out/r2p_ncu.log 2>&1; echo "ncu list rc=$?"
