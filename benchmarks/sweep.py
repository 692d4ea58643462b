#!/usr/bin/env python
"""BASELINE.json configs[3]: standalone BN254 NTT + G1/G2 MSM sweep (default 2^16..2^22; --max-log 24 on a free box).

    python benchmarks/sweep.py [--min-log 16] [--max-log 22] [--cpu-max-log 18] > profiles/sweep.jsonl

Per size: GPU kernel time (CUDA events inside the library; MSM time excludes the one-time window-table expansion,
reported separately as table_ms... see nzcp_msm), achieved Fq/Fr mul/s and HBM GB/s against the measured peaks, and
the host-CPU time of the C restatement of ffjavascript's algorithms (oracle/c, all threads) up to --cpu-max-log.
Uniform 254-bit scalars, bases k_i * G for pseudo-random k_i.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nzcp_circom_b200 import api  # noqa: E402
from oracle import cref  # noqa: E402

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def rand_fr(n, seed):
    rs = np.random.RandomState(seed)
    x = rs.randint(0, 2 ** 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    x[:, 7] &= 0x1FFFFFFF          # < 2^253 < r
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log", type=int, default=16)
    ap.add_argument("--max-log", type=int, default=22)
    ap.add_argument("--cpu-max-log", type=int, default=18)
    ap.add_argument("--g2-max-log", type=int, default=20)
    args = ap.parse_args()
    ip = api.intpipe_bench(0, 4096)
    peak_mul = ip["imad_wide_per_s"] / 128
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    threads = cref.max_threads()
    for lg in range(args.min_log, args.max_log + 1):
        n = 1 << lg
        # ---- NTT (forward, natural order in and out)
        x = rand_fr(n, lg)
        api.ntt(x.copy(), lg)                                   # warm-up (tables, module load)
        ms = min(api.ntt(x.copy(), lg) for _ in range(3))
        muls = (n // 2) * lg
        rec = {"op": "ntt", "log_n": lg, "gpu_ms": ms, "GFrmul_s": muls / ms / 1e6, "frac_int_pipe": muls / (ms * 1e-3) / peak_mul,
               "GB_s": 2 * 32 * n / ms / 1e6, "frac_hbm": 2 * 32 * n / ms / 1e6 / hbm}
        if lg <= args.cpu_max_log:
            y = x.copy()
            t = time.perf_counter()
            cref.ntt(y, lg, False, threads)
            rec["cpu_ms"] = 1e3 * (time.perf_counter() - t)
            rec["cpu_threads"] = threads
            z = x.copy()
            api.ntt(z, lg)
            rec["matches_cpu_port"] = bool(np.array_equal(y, z))
        print(json.dumps(rec), flush=True)
        # ---- MSM G1 / G2, uniform scalars
        sc = rand_fr(n, 1000 + lg)
        for g2 in (False, True):
            if g2 and lg > args.g2_max_log:
                continue
            bases = api.synth_points(77 + lg, n, g2=g2)
            api.msm(bases, sc, n, g2=g2)
            out, ms = None, 1e30
            for _ in range(2):
                o, m = api.msm(bases, sc, n, g2=g2)
                out, ms = o, min(ms, m)
            c = 16 if lg >= 19 else (14 if lg >= 17 else 12)
            windows = -(-255 // c)
            muls = n * windows * (28 if g2 else 10)
            byts = n * ((128 if g2 else 64) + 32)
            rec = {"op": "msm_g2" if g2 else "msm_g1", "log_n": lg, "gpu_ms": ms, "GFqmul_s": muls / ms / 1e6,
                   "frac_int_pipe": muls / (ms * 1e-3) / peak_mul, "Mpoints_s": n / ms / 1e3,
                   "alg_GB_s": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / hbm,
                   "note": "gpu_ms = sort + accumulate + combine + reduce; window-table expansion (once per base set) excluded"}
            if lg <= args.cpu_max_log:
                t = time.perf_counter()
                ref = cref.msm(bases, sc, n, g2, threads)
                rec["cpu_ms"] = 1e3 * (time.perf_counter() - t)
                rec["cpu_threads"] = threads
                rec["matches_cpu_port"] = (ref == out)
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
