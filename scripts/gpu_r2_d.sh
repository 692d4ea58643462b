#!/bin/bash
# round 2, call D: H-window sweep, pair rounds re-check, launch list + ncu --set full on the final default, NTT/MSM sweep
mkdir -p gpurun_out
run() { # tag, args...
  tag=$1; shift
  timeout 400 python bench.py --no-extras --no-cpu-baseline "$@" > gpurun_out/r2d_bench_$tag.json 2> gpurun_out/r2d_bench_$tag.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2d_bench_$tag.json"))
    print("$tag rc=$rc value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "p50", round(d["p50_latency_ms"],2), "acc_h", round(d["msm"]["accumulate_ms"]["h"],3), "sort_h", round(d["msm"]["sort_ms"]["h"],3), "msm_h", round(d["stage_ms"]["msm_h"],3), "entries_h", d["msm"]["n_entries"]["h"])
except Exception as e:
    print("$tag rc=$rc parse failed", e)
PY
}
run default
run ch17 --tune prover_c_h=17
run ch18 --tune prover_c_h=18
run ch20 --tune prover_c_h=20
run cw14 --tune prover_c_w=14
run rounds_h3 --tune prover_rounds_h=3
run rounds_h3w2 --tune prover_rounds_h=3 --tune prover_rounds_w=2
run rounds_h2 --tune prover_rounds_h=2
# launch list of one proof on the final default build
python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras > gpurun_out/r2d_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 700 --csv --log-file gpurun_out/r2d_launches_pipe.csv python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras > gpurun_out/r2d_ncu.log 2>&1; echo "ncu list rc=$?"
# full capture: the NTT kernels (incl. the TMA-staged low pass) and the accumulate kernels of one proof
ncu --set full --import-source on --clock-control none -k regex:"ntt_|msm_accumulate_kernel" --launch-skip 32 --launch-count 8 -o gpurun_out/r2d_prof_ntt_acc -f python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline --no-extras > gpurun_out/r2d_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2d_prof_ntt_acc.ncu-rep --page raw --csv > gpurun_out/r2d_prof_ntt_acc_raw.csv 2>/dev/null; echo "raw rc=$?"
timeout 900 python benchmarks/sweep.py --min-log 16 --max-log 24 --g2-max-log 24 --cpu-max-log 20 > gpurun_out/r2d_sweep.jsonl 2> gpurun_out/r2d_sweep.err; echo "sweep rc=$?"; tail -2 gpurun_out/r2d_sweep.err
python - <<'PY'
import json
for l in open("gpurun_out/r2d_sweep.jsonl"):
    d = json.loads(l)
    print(d["op"], d.get("mode", "")[:8], d["log_n"], "gpu_ms", round(d["gpu_ms"], 3), "table", round(d.get("table_ms", 0), 1), "wall", round(d.get("wall_ms", 0), 2), "frac", round(d.get("frac_int_pipe", 0), 3), "cpu_ms", round(d.get("cpu_ms", 0), 1), d.get("matches_cpu_port"), d.get("same_point_as_fixed_base"))
PY
