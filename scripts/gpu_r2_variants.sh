#!/bin/bash
# A/B of pair-round settings through bench.py (one line per variant), after a quick parity run of the MSM paths.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "pair_rounds or msm or benchmarked_shape or forced" > gpurun_out/r2_pytest_msm.log 2>&1; tail -3 gpurun_out/r2_pytest_msm.log
run() {
  name=$1; shift
  envs=$1; shift
  env $envs python bench.py --no-cpu-baseline "$@" > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/v_$name.json"))
    print("%-12s"%"$name", "value %.2f"%d["value"], "e2e %.2f"%d["e2e"]["value"], "p50 %.2f"%d["p50_latency_ms"], "acc", {k:round(v,2) for k,v in d["msm"]["accumulate_ms"].items()}, "ntt %.2f"%d["stage_ms"]["ntt_join"])
except Exception as e:
    print("$name failed", e)
PY
}
run base "X=1" --tune prover_rounds_w=0 --tune prover_rounds_h=0
run h3k16 "X=1" --tune prover_rounds_w=0
run h3k8 "X=1" --tune prover_rounds_w=0 --tune pair_k1=8 --tune pair_k2=8 --tune pair_k3=8
run h3k32 "X=1" --tune prover_rounds_w=0 --tune pair_k1=32 --tune pair_k2=32 --tune pair_k3=32
run h2k16 "X=1" --tune prover_rounds_w=0 --tune prover_rounds_h=2
run w2h3k16 "X=1"
run w2h3k8 "X=1" --tune pair_k1=8 --tune pair_k2=8 --tune pair_k3=8
run w1h3k8 "X=1" --tune prover_rounds_w=1 --tune pair_k1=8 --tune pair_k2=8 --tune pair_k3=8
KERNEL="msm_pair_|msm_accumulate_pts" SKIP=${SKIP:-136} COUNT=${COUNT:-50} TAG=pair3 TUNE="" bash scripts/gpu_r2_prof.sh > gpurun_out/prof3.log 2>&1; tail -3 gpurun_out/prof3.log
