#!/bin/bash
# round 2, call F2 (2 GPUs): two-device test, split MSM through NCCL checked against the C oracle, bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_prove.py -m gpu -q -k two_devices > gpurun_out/r2f2_pytest.log 2>&1; echo "two-device test rc=$?"; tail -3 gpurun_out/r2f2_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $T benchmarks/msm_split.py --log-n 20 --check > gpurun_out/r2f2_split20.json 2> gpurun_out/r2f2_split20.err; echo "split 2^20 rc=$?"; cat gpurun_out/r2f2_split20.json | cut -c1-700
timeout 600 $T benchmarks/msm_split.py --log-n 20 --mode 0 --check > gpurun_out/r2f2_split20_fixed.json 2> gpurun_out/r2f2_split20_fixed.err; echo "split 2^20 fixed rc=$?"; cat gpurun_out/r2f2_split20_fixed.json | cut -c1-400
timeout 600 $T benchmarks/msm_split.py --log-n 24 > gpurun_out/r2f2_split24.json 2> gpurun_out/r2f2_split24.err; echo "split 2^24 rc=$?"; cat gpurun_out/r2f2_split24.json | cut -c1-500
( time timeout 900 $T bench.py --gpus 2 ) > gpurun_out/r2f2_bench_n2.json 2> gpurun_out/r2f2_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2f2_bench_n2.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2f2_bench_n2.json"))
    print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "p50", round(d["p50_latency_ms"],2), "launches", d["gpu_launches"], d.get("proof_verifies"), d.get("proof_matches_cpu_port"))
    print(json.dumps(d.get("extras"))[:1500])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 $T bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2f2_ref_n2.json 2> gpurun_out/r2f2_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-200 gpurun_out/r2f2_ref_n2.json
