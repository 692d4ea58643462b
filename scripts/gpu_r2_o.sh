#!/bin/bash
# round 2, call O: interleaved pair assignment -- parity, then the bench (compare with 129.8-131.5 of the blocked assignment)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_prove.py -m gpu -q -k "pair or benchmarked or rounds or prove_matches or msm_dense" > gpurun_out/r2o_pytest.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2o_pytest.log
for rep in 1 2; do
timeout 400 python bench.py --no-extras --no-cpu-baseline --latency-runs 20 > gpurun_out/r2o_bench_$rep.json 2> gpurun_out/r2o_bench_$rep.err; rc=$?
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2o_bench_$rep.json"))
    r=d["roofline"]
    print("rep $rep rc=$rc value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2), d["latency_ms"]["min"], "roofline", round(r["frac"],3), r["launch_ms"], r["in_proof_ms"], (r.get("xyzz_kernel") or {}).get("launch_ms"), "step", round(d["roofline_step"]["frac"],3))
except Exception as e:
    print("rep $rep rc=$rc parse failed", e)
PY
done
