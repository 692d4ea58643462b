# 2-GPU checks: bench.py under torchrun (batch sharding, weak scaling), the reference arm, the split MSM
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 3 --warmup 3 --batch 6 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo bench rc=$?; cat gpurun_out/bench_n$N.json | cut -c1-600; tail -3 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29502 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/ref_n$N.json 2> gpurun_out/ref_n$N.err; echo ref rc=$?; cat gpurun_out/ref_n$N.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29503 benchmarks/msm_split.py --log-n ${LOGN:-22} > gpurun_out/msm_split_n$N.json 2> gpurun_out/msm_split_n$N.err; echo split rc=$?; cat gpurun_out/msm_split_n$N.json; tail -3 gpurun_out/msm_split_n$N.err
python benchmarks/msm_split.py --log-n ${LOGN:-22} > gpurun_out/msm_split_n1.json 2> gpurun_out/msm_split_n1.err; echo split1 rc=$?; cat gpurun_out/msm_split_n1.json
