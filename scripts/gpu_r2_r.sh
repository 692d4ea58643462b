#!/bin/bash
# round 2, call R (1 GPU): the build that is committed last -- whole GPU suite, smoke, reference arm, bench of record,
# and the XYZZ accumulate kernel's DRAM traffic with / without the gather fetch-size qualifier (standalone fixed-base MSM)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2r_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r2r_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2r_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_ref.json 2> gpurun_out/r2r_ref.err; echo "ref rc=$?"
( time python bench.py ) > gpurun_out/r2r_bench_default.json 2> gpurun_out/r2r_bench_default.err; echo "bench rc=$?"; tail -4 gpurun_out/r2r_bench_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2r_bench_default.json"))
    r=d["roofline"]
    print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2), round(d["p50_latency_ms_pageable"],2), "launches", d["gpu_launches"], d.get("proof_verifies"), d.get("proof_matches_cpu_port"))
    print("roofline", round(r["frac"],3), r["launch_ms"], r["in_proof_ms"], (r.get("xyzz_kernel") or {}).get("launch_ms"), "step", round(d["roofline_step"]["frac"],3), round(d["roofline_step"]["frac_without_reduction_term"],3))
    rr=json.load(open("gpurun_out/r2r_ref.json")); print("ref", rr["value"], rr["cpu_baseline"]["cores"], rr["config"]==d["config"])
except Exception as e:
    print("bench parse failed", e)
PY
cat > /tmp/xyzz_hint.py <<'PY'
import sys, numpy as np
from nzcp_circom_b200 import api
n = 1 << 20
bases = bytes(api.synth_points(77, n))
sc = np.random.RandomState(5).randint(0, 2 ** 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); sc[:, 7] &= 0x1FFFFFFF
api.tuning_set("gather_hint", int(sys.argv[1]))
with api.MsmPlan(bases, n, mode=0) as plan:
    for _ in range(3):
        plan.run(sc)
    print("gather_hint", sys.argv[1], "accumulate ms", plan.accumulate_ms())
PY
for h in 0 1; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum -k regex:"msm_accumulate_kernel" --clock-control none --launch-skip 2 --launch-count 1 --csv --log-file gpurun_out/r2r_xyzz_hint$h.csv python /tmp/xyzz_hint.py $h > gpurun_out/r2r_xyzz_hint$h.log 2>&1
grep -E "msm_accumulate" gpurun_out/r2r_xyzz_hint$h.csv | awk -F'","' '{print "hint='$h'", $(NF-2), $(NF-1), $NF}'
python /tmp/xyzz_hint.py $h
done
