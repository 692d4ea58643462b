#!/bin/bash
# round 2, call Q: ncu --set full of the final build's pair-round kernels + XYZZ tail (one proof's worth of launches);
# only the raw-page CSV travels back (the .ncu-rep of 20 launches with sources exceeds the 64 MiB return limit)
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"msm_pair_|msm_accumulate_pts" --launch-skip 200 --launch-count 20 -o /tmp/r2q_prof_pair -f python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras --latency-runs 2 > gpurun_out/r2q_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r2q_prof_pair.ncu-rep --page raw --csv > gpurun_out/r2q_prof_pair_raw.csv 2>/dev/null; echo "raw rc=$?"
ls -la /tmp/r2q_prof_pair.ncu-rep gpurun_out/r2q_prof_pair_raw.csv
# experiment: .L2::64B fetch-size qualifier on the round-1 gathers (DRAM bytes per launch + time), then the bench both ways
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum -k regex:"msm_pair_(forward|backward)_kernel" --clock-control none --launch-skip 200 --launch-count 12 --csv --log-file gpurun_out/r2q_hint0.csv python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras --latency-runs 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum -k regex:"msm_pair_(forward|backward)_kernel" --clock-control none --launch-skip 200 --launch-count 12 --csv --log-file gpurun_out/r2q_hint1.csv python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras --latency-runs 2 --tune gather_hint=1 > /dev/null 2>&1
python - <<'PY'
import csv
for tag in ("hint0", "hint1"):
    rows = list(csv.DictReader([l for l in open("gpurun_out/r2q_%s.csv" % tag) if not l.startswith("==")]))
    by = {}
    for r in rows:
        by.setdefault(r["ID"], {"name": r["Kernel Name"][:60], "grid": r["Grid Size"]})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    for k, v in by.items():
        print(tag, v["name"], v["grid"], v.get("gpu__time_duration.sum"), v.get("dram__bytes_read.sum"))
PY
for h in 0 1; do
timeout 300 python bench.py --no-extras --no-cpu-baseline --latency-runs 10 --tune gather_hint=$h > gpurun_out/r2q_bench_hint$h.json 2> gpurun_out/r2q_bench_hint$h.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2q_bench_hint$h.json")); print("hint=$h value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "p50", round(d["p50_latency_ms"],2), "H alone", d["roofline"]["launch_ms"])
except Exception as e:
    print("hint=$h failed", e)
PY
done
