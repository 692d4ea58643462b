#!/bin/bash
# round 2, call F1 (1 GPU): prover modes -- full-size parity in both modes, then the bench of record + reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_prove.py tests/test_napi_shim.py -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2f_pytest.log
( time python bench.py ) > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; echo "bench rc=$?"; tail -4 gpurun_out/r2f_bench_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2f_bench_default.json"))
    print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), round(d["e2e"]["pageable"]["value_driver_staged"],2), "p50", round(d["p50_latency_ms"],2), round(d["p50_latency_ms_pageable"],2), "launches", d["gpu_launches"], "verifies", d.get("proof_verifies"), d.get("proof_matches_cpu_port"))
    print("roofline", round(d["roofline"]["frac"],3), d["roofline"]["launch_ms"], "step", json.dumps(d["roofline_step"])[:900])
    print(json.dumps(d.get("extras"))[:1500])
except Exception as e:
    print("bench parse failed", e)
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2f_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2f_smoke.log
