# parity suite + launch list (kernel times) + DRAM bytes of the accumulate kernels under different L2 fetch granularities
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
bash scripts/gpu_launches.sh
python scripts/launch_summary.py gpurun_out/launches.csv | head -12
CMD="python bench.py --steps 1 --warmup 3 --batch 1 --provers 1 --no-cpu-baseline"
for G in 32 64 128; do
  NZCP_L2_FETCH=$G ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:msm_accumulate_kernel -s 20 -c 5 --csv --log-file gpurun_out/l2fetch_$G.csv $CMD > gpurun_out/l2fetch_$G.log 2>&1; echo "G=$G rc=$?"
  grep -E "dram__bytes_read|gpu__time" gpurun_out/l2fetch_$G.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | sed 's/"//g'
done
