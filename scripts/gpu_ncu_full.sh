# one `ncu --set full` capture of the accumulate kernels (G2 witness, 3x G1 witness, G1 H) of one proof
CMD="python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel -s 20 -c 5 -o gpurun_out/prof_accumulate $CMD > gpurun_out/ncu_full.log 2>&1; echo ncu rc=$?
