# end-of-round checks: smoke(), the reference arm, and an ncu --set full capture of the memory-side kernels
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null > gpurun_out/ref_n1.json; cut -c1-400 gpurun_out/ref_n1.json
KERNEL="r1cs_eval|join_abc|msm_digits_smem" SKIP=12 COUNT=6 TAG=mem bash scripts/gpu_ncu_full.sh
ncu -i gpurun_out/prof_mem.ncu-rep --page raw --csv > gpurun_out/prof_mem_raw.csv 2>/dev/null; echo raw rc=$?
