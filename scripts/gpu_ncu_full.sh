# `ncu --set full` captures: KERNEL regex, SKIP launches, COUNT launches -> gpurun_out/prof_$TAG.ncu-rep
CMD="python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-msm_accumulate_kernel} -s ${SKIP:-20} -c ${COUNT:-5} -o gpurun_out/prof_${TAG:-accumulate} $CMD > gpurun_out/ncu_full.log 2>&1; echo ncu rc=$?
