#!/usr/bin/env python
"""BASELINE.json configs[4]: ONE 2^log_n-point G1 MSM split over the GPUs of a box by point range, NCCL gather of the
partial sums.  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
                        --master-port 29511 benchmarks/msm_split.py --log-n 24 [--mode 1] [--check]

Product path (nzcp_circom_b200.parallel.msm_split does the same from full arrays): every rank generates only ITS slice
of the bases (k_i * G, seed = rank) and scalars, keeps the bases resident in an MSM plan (mode 1 = variable-base, no
window table; mode 0 = fixed-base table), runs the plan -> ONE XYZZ point left in HBM, all-gathers the points over
NCCL (device buffers on both sides), and a kernel adds them.  Timed per repetition with CUDA events between barriers,
max over ranks; the plan build (bases upload, table expansion in mode 0) is reported separately.  --check compares the
result with the C oracle on rank 0 (sizes the CPU finishes in seconds).  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nzcp_circom_b200 import api, parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=22)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--mode", type=int, default=1, choices=[0, 1])
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << args.log_n
    lo, hi = parallel.shard_range(n, rank, world)
    cnt = hi - lo
    bases = api.synth_points(1000 + rank, cnt, device=local)
    rs = np.random.RandomState(rank)
    sc = rs.randint(0, 2 ** 32, size=(cnt, 8), dtype=np.uint64).astype(np.uint32)
    sc[:, 7] &= 0x1FFFFFFF
    mine = torch.zeros(128, dtype=torch.uint8, device=dev)
    kernel_ms, step_ms = [], []
    result = None
    with api.MsmPlan(bases, cnt, mode=args.mode, device=local) as plan:
        for rep in range(args.reps + 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            kms = plan.run_partial(sc, mine)                      # scalars H2D + sort + accumulate + reduce + fold
            if world > 1:
                parts = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine)                      # 128 B per rank, device to device over NVLink
            else:
                parts = [mine]
            result = api.msm_sum_partials(torch.cat(parts), world, device=local)
            e1.record()
            torch.cuda.synchronize()
            if rep:                                               # first repetition is warm-up
                kernel_ms.append(kms)
                step_ms.append(e0.elapsed_time(e1))
        build_ms = plan.build_ms
    t = torch.tensor([min(kernel_ms), min(step_ms), build_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    check = None
    if args.check:
        # every rank's slice is reproducible from its seed: rank 0 rebuilds the whole problem for the CPU oracle
        if rank == 0:
            from oracle import cref
            allb, alls = [], []
            for r in range(world):
                l2, h2 = parallel.shard_range(n, r, world)
                allb.append(bytes(api.synth_points(1000 + r, h2 - l2, device=local)))
                s2 = np.random.RandomState(r).randint(0, 2 ** 32, size=(h2 - l2, 8), dtype=np.uint64).astype(np.uint32)
                s2[:, 7] &= 0x1FFFFFFF
                alls.append(s2.tobytes())
            check = cref.msm(b"".join(allb), b"".join(alls), n, False, cref.max_threads()) == result
    if rank == 0:
        k, s, b = (float(x) for x in t.tolist())
        print(json.dumps({"op": "msm_g1_split", "log_n": args.log_n, "n_gpus": world, "points_per_gpu": cnt,
                          "mode": "variable-base (no table)" if args.mode == 1 else "fixed-base (window table)",
                          "kernel_ms_max_over_ranks": k, "step_ms_max_over_ranks": s, "plan_build_ms_max_over_ranks": b,
                          "Mpoints_s_kernel": n / k / 1e3, "Mpoints_s_step": n / s / 1e3,
                          "result_x": str(int.from_bytes(result[:32], "little")), "matches_cpu_port": check,
                          "note": "step = scalar upload from pageable host memory + MSM kernels + NCCL all-gather of one XYZZ "
                                  "point per rank + sum kernel + 64-byte result to host; bases resident (plan build separate)"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
