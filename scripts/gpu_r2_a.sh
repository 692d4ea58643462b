#!/bin/bash
# round 2, call A: new-entry-point tests first (fast feedback), then the full GPU suite, the default bench, and a
# per-kernel launch list of one proof carrying the multiplier-pipe activity next to the duration (whole-step pipe time).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm_plan.py -m gpu -x -q > gpurun_out/r2a_pytest_new.log 2>&1; echo "new tests rc=$?"; tail -15 gpurun_out/r2a_pytest_new.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_msm_plan.py > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -8 gpurun_out/r2a_pytest_gpu.log
python bench.py > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2a_bench_default.json"))
    print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "p50", round(d["p50_latency_ms"],2), "launches", d["gpu_launches"], "acc", {k:round(v,3) for k,v in d["msm"]["accumulate_ms"].items()}, "frac", round(d["roofline"]["frac"],3), d.get("proof_matches_cpu_port"))
    print(d["stage_ms"])
except Exception as e:
    print("bench parse failed", e)
PY
ncu --metrics gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches_pipe.csv python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2a_ncu.log
