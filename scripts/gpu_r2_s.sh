#!/bin/bash
mkdir -p gpurun_out
for h in 0 1; do
python scripts/xyzz_gather_hint.py $h
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum -k regex:"msm_accumulate_kernel" --clock-control none --launch-skip 2 --launch-count 1 --csv --log-file gpurun_out/r2s_xyzz_hint$h.csv python scripts/xyzz_gather_hint.py $h > gpurun_out/r2s_xyzz_hint$h.log 2>&1
grep -E "msm_accumulate" gpurun_out/r2s_xyzz_hint$h.csv | awk -F'","' '{print "hint='$h'", $(NF-2), $(NF-1), $NF}'
done
