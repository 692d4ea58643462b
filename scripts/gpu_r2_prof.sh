#!/bin/bash
# launch list of one proof + ncu --set full of the pair-round kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline ${TUNE}"
$CMD > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1; echo ncu list rc=$?
python scripts/launch_summary.py gpurun_out/launches.csv -v > gpurun_out/launch_summary.txt 2>&1; head -40 gpurun_out/launch_summary.txt
ncu --set full --clock-control none --import-source on -k regex:"${KERNEL:-msm_pair_round|msm_accumulate_pts}" -s ${SKIP:-48} -c ${COUNT:-16} -o gpurun_out/prof_${TAG:-pair} $CMD > gpurun_out/ncu_full.log 2>&1; echo ncu full rc=$?
ncu -i gpurun_out/prof_${TAG:-pair}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG:-pair}_raw.csv 2>/dev/null
python scripts/ncu_raw_to_md.py gpurun_out/prof_${TAG:-pair}_raw.csv > gpurun_out/prof_${TAG:-pair}.md; wc -l gpurun_out/prof_${TAG:-pair}.md
[ -n "$KEEP_REP" ] || rm -f gpurun_out/prof_${TAG:-pair}.ncu-rep
