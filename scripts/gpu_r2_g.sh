#!/bin/bash
# round 2, call G (N GPUs, N = number visible): bench at N, split MSM 2^24 at N
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
( time timeout 900 $T bench.py --gpus $N ) > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/r2g_bench_n$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2g_bench_n$N.json"))
    print("n_gpus", d["n_gpus"], "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "p50", round(d["p50_latency_ms"],2), d.get("proof_verifies"), d.get("proof_matches_cpu_port"), d["clocks"])
    print(json.dumps(d.get("extras"))[:1600])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 $T benchmarks/msm_split.py --log-n 24 > gpurun_out/r2g_split24_n$N.json 2> gpurun_out/r2g_split24_n$N.err; echo "split 2^24 rc=$?"; cut -c1-600 gpurun_out/r2g_split24_n$N.json
timeout 600 $T benchmarks/msm_split.py --log-n 22 --check > gpurun_out/r2g_split22_n$N.json 2> gpurun_out/r2g_split22_n$N.err; echo "split 2^22 check rc=$?"; cut -c1-600 gpurun_out/r2g_split22_n$N.json
