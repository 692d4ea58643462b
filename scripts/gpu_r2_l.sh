#!/bin/bash
# round 2, call L: operand staging in round 1 of the pair rounds -- parity, then A/B against gathering twice
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_prove.py -m gpu -q -k "pair or benchmarked or rounds or prove_matches" > gpurun_out/r2l_pytest.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2l_pytest.log
run() { # tag, args...
  tag=$1; shift
  timeout 400 python bench.py --no-extras --no-cpu-baseline --latency-runs 20 "$@" > gpurun_out/r2l_bench_$tag.json 2> gpurun_out/r2l_bench_$tag.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2l_bench_$tag.json"))
    print("$tag rc=$rc value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), "p50", round(d["p50_latency_ms"],2), "acc_h", round(d["msm"]["accumulate_ms"]["h"],2))
except Exception as e:
    print("$tag rc=$rc parse failed", e)
PY
}
run stage1
run stage0 --tune pair_stage=0
run stage1_b
run stage0_b --tune pair_stage=0
run stage1_k32 --tune pair_k1=32
run stage1_k8 --tune pair_k1=8
python - <<'PY'
import numpy as np
from nzcp_circom_b200 import api
n = 1 << 20
bases = bytes(api.synth_points(77, n))
rs = np.random.RandomState(5)
sc = rs.randint(0, 2 ** 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); sc[:, 7] &= 0x1FFFFFFF
for rounds in (0, 3, 0, 3):
    api.tuning_set("msm_rounds", rounds)
    with api.MsmPlan(bases, n, mode=0) as plan:
        plan.run(sc)
        ms = min(plan.run(sc)[1] for _ in range(4))
    print("standalone fixed-base G1 2^20 dense, rounds", rounds, "kernel ms", round(ms, 3))
PY
