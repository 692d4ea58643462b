#!/bin/bash
# round 2, first GPU call: full GPU suite (incl. the new full-size parity tests and the forced pair rounds), then the
# bench with and without the batched-affine pair rounds.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
python bench.py --no-cpu-baseline --tune prover_rounds_h=0 --tune prover_rounds_w=0 > gpurun_out/r2_bench_norounds.json 2> gpurun_out/r2_bench_norounds.err
python bench.py --no-cpu-baseline --tune prover_rounds_w=0 > gpurun_out/r2_bench_honly.json 2> gpurun_out/r2_bench_honly.err
python bench.py --no-cpu-baseline --tune pair_k1=64 --tune pair_k2=64 --tune pair_k3=32 > gpurun_out/r2_bench_k64.json 2> gpurun_out/r2_bench_k64.err
for f in default norounds honly k64; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_$f.json"))
    print("$f", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "p50", round(d["p50_latency_ms"],2), "acc", {k:round(v,3) for k,v in d["msm"]["accumulate_ms"].items()}, "frac", round(d["roofline"]["frac"],3), d.get("proof_matches_cpu_port"))
except Exception as e:
    print("$f failed", e)
PY
done
