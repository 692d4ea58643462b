#!/bin/bash
# round 2, call C: zkey-new tests, the TMA-staged NTT low pass (parity first, then timing both ways), threaded staging
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zkey_new.py "tests/test_gpu_prove.py::test_groth16_prove_thread_safety_and_bounded_cache" tests/test_gpu_kernels.py -m gpu -q -k "zkey_new or thread_safety or ntt" > gpurun_out/r2c_pytest_new.log 2>&1; echo "new tests rc=$?"; tail -25 gpurun_out/r2c_pytest_new.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "ntt or benchmarked" > gpurun_out/r2c_pytest_full.log 2>&1; echo "fullsize rc=$?"; tail -5 gpurun_out/r2c_pytest_full.log
timeout 300 python - <<'PY'
import numpy as np
from nzcp_circom_b200 import api
for log_n in (20, 22):
    n = 1 << log_n
    buf = np.random.default_rng(1).integers(0, 256, size=3 * n * 32, dtype=np.uint8)
    buf[31::32] &= 0x1f
    for tma in (1, 0, 1, 0):
        api.tuning_set("ntt_tma", tma)
        api.ntt_coset(buf, log_n, batch=3)
        ms = min(api.ntt_coset(buf, log_n, batch=3) for _ in range(5))
        print("ntt_coset batch 3 log_n", log_n, "tma", tma, "ms", round(ms, 4))
PY
for t in 1 0; do
timeout 300 python bench.py --no-extras --no-cpu-baseline --tune ntt_tma=$t > gpurun_out/r2c_bench_tma$t.json 2> gpurun_out/r2c_bench_tma$t.err; echo "bench tma=$t rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2c_bench_tma$t.json"))
    print("tma=$t value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "pageable", round(d["e2e"]["pageable"]["value"],2), round(d["e2e"]["pageable"]["value_driver_staged"],2), "p50", round(d["p50_latency_ms"],2), round(d["p50_latency_ms_pageable"],2), "ntt_join", round(d["stage_ms"]["ntt_join"],3), "ntt alone", d["roofline_ntt"]["alone"]["ms"])
except Exception as e:
    print("bench parse failed", e)
PY
done
