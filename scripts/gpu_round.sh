# One validation round: GPU parity suite, default bench, ncu launch list, ncu --set full of the accumulate + NTT kernels
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
python bench.py ${BENCH_ARGS} > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'p50',round(d['p50_latency_ms'],2), 'launches', d['gpu_launches'])
print(d['stage_ms']); print(d['msm']); print(d['roofline']['frac'], d['roofline']['launch_ms']); print(d['roofline_ntt'].get('alone')); print(d.get('proof_matches_cpu_port'))
PY
if [ -n "$NCU" ]; then
bash scripts/gpu_launches.sh
python scripts/launch_summary.py gpurun_out/launches.csv | head -30
KERNEL="msm_accumulate_kernel|ntt_" SKIP=32 COUNT=8 TAG=acc_ntt bash scripts/gpu_ncu_full.sh
ncu -i gpurun_out/prof_acc_ntt.ncu-rep --page raw --csv > gpurun_out/prof_acc_ntt_raw.csv 2>/dev/null; echo raw rc=$?
fi
