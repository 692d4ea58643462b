#!/bin/bash
# round 2, call U: last validation of the committed build -- whole GPU suite + smoke + a short bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r2u_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2u_smoke.log
python bench.py --no-extras > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2u_bench.json"))
print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "p50", round(d["p50_latency_ms"],2), d.get("proof_verifies"), d.get("proof_matches_cpu_port"), "roofline", round(d["roofline"]["frac"],3))
PY
