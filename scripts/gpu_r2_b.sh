#!/bin/bash
# round 2, call B: the new tests (MSM plans incl. the device-finish fix, N-API shim through the mock runtime, groth16 cache),
# then the bench with extras, the reference arm, and a smaller-extras 1-GPU sanity of the collective-free path.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm_plan.py tests/test_napi_shim.py "tests/test_gpu_prove.py::test_groth16_prove_thread_safety_and_bounded_cache" tests/test_gpu_prove.py::test_snarkjs_surface_and_verify -m gpu -q > gpurun_out/r2b_pytest_new.log 2>&1; echo "new tests rc=$?"; tail -25 gpurun_out/r2b_pytest_new.log
( time python bench.py ) > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err; echo "bench rc=$?"; tail -4 gpurun_out/r2b_bench_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_bench_default.json"))
    print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", d["e2e"]["pageable"]["value"], d["e2e"]["pageable"]["value_driver_staged"], "p50", round(d["p50_latency_ms"],2), d["p50_latency_ms_pageable"], "launches", d["gpu_launches"], "verifies", d.get("proof_verifies"), d.get("proof_matches_cpu_port"))
    print(json.dumps(d.get("extras"), indent=1)[:3000])
except Exception as e:
    print("bench parse failed", e)
PY
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2b_ref.json 2> gpurun_out/r2b_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/r2b_ref.err; cut -c1-600 gpurun_out/r2b_ref.json
