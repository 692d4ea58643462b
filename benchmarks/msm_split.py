#!/usr/bin/env python
"""BASELINE.json configs[4]: ONE 2^log_n-point G1 MSM split over the GPUs of a box by point range, NCCL gather of the
partial sums.  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
                        --master-port 29511 benchmarks/msm_split.py --log-n 24
Every rank generates only ITS slice of the bases (k_i * G, seed = slice index) and scalars, runs the single-GPU MSM on
it, then one all-gather of 64 bytes per rank and N-1 additions.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nzcp_circom_b200 import api, parallel, verifier  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=22)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << args.log_n
    lo, hi = parallel.shard_range(n, rank, world)
    cnt = hi - lo
    bases = api.synth_points(1000 + rank, cnt, device=local)
    rs = np.random.RandomState(rank)
    sc = rs.randint(0, 2 ** 32, size=(cnt, 8), dtype=np.uint64).astype(np.uint32)
    sc[:, 7] &= 0x1FFFFFFF
    times, kernel_ms = [], []
    for rep in range(args.reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        part, kms = api.msm(bases, sc, cnt, device=local)
        mine = torch.frombuffer(bytearray(part), dtype=torch.uint8).cuda()
        if world > 1:
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
        else:
            parts = [mine]
        acc = None
        for p in parts:
            acc = verifier.g1_add(acc, parallel._point_from_bytes(bytes(p.cpu().numpy()), False))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if rep:                      # first repetition is warm-up
            times.append(dt)
            kernel_ms.append(kms)
    k = torch.tensor([min(kernel_ms)], device="cuda")
    if world > 1:
        dist.all_reduce(k, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"op": "msm_g1_split", "log_n": args.log_n, "n_gpus": world, "points_per_gpu": cnt,
                          "kernel_ms_max_over_ranks": float(k.item()),
                          "Mpoints_s": n / float(k.item()) / 1e3,
                          "wall_ms_incl_h2d_table_gather": 1e3 * min(times),
                          "result_x": str(acc[0]) if acc else None,
                          "note": "kernel_ms = sort+accumulate+combine+reduce on the slowest rank; wall includes the host->"
                                  "device copy of bases/scalars and the one-time window-table expansion of nzcp_msm"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
