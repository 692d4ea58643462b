#!/usr/bin/env python
"""BASELINE.json configs[3]: standalone BN254 NTT + G1/G2 MSM sweep (default 2^16..2^22; --max-log 24 on a free box).

    python benchmarks/sweep.py [--min-log 16] [--max-log 22] [--cpu-max-log 18] > profiles/sweep.jsonl

Per size: GPU kernel time (CUDA events inside the library), achieved Fq/Fr mul/s and HBM GB/s against the measured
peaks, and the host-CPU time of the C restatement of ffjavascript's algorithms (oracle/c, all threads) up to
--cpu-max-log, with the results compared.  Every MSM is run in both modes: fixed-base (window table; table_ms reported
next to the per-MSM gpu_ms) and variable-base (no table; wall_ms of the whole one-shot call next to gpu_ms).
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
    ap.add_argument("--cpu-max-log", type=int, default=20)
    ap.add_argument("--g2-max-log", type=int, default=22)
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
        # ---- MSM G1 / G2, uniform scalars: both modes of the library
        #   fixed-base   bases expanded once into the 2^(c w) window table (what the prover does per zkey section):
        #                table_ms is the one-time build, gpu_ms each MSM after it
        #   variable-base one-shot multiExpAffine semantics, no table: wall_ms covers uploads + plan + kernels + result
        sc = rand_fr(n, 1000 + lg)
        for g2 in (False, True):
            if g2 and lg > args.g2_max_log:
                continue
            bases = api.synth_points(77 + lg, n, g2=g2)
            per_add = 28 if g2 else 10
            byts = n * ((128 if g2 else 64) + 32)
            ref = None
            if lg <= args.cpu_max_log:
                t = time.perf_counter()
                ref = cref.msm(bases, sc, n, g2, threads)
                cpu_ms = 1e3 * (time.perf_counter() - t)
            with api.MsmPlan(bases, n, g2=g2, mode=0) as plan:
                plan.run(sc)
                out, ms = None, 1e30
                for _ in range(2):
                    o, m = plan.run(sc)
                    out, ms = o, min(ms, m)
                table_ms = plan.build_ms
            c = 16 if lg >= 19 else (14 if lg >= 17 else 12)
            muls = n * -(-255 // c) * per_add
            rec = {"op": "msm_g2" if g2 else "msm_g1", "mode": "fixed-base (window table)", "log_n": lg, "gpu_ms": ms,
                   "table_ms": table_ms, "GFqmul_s": muls / ms / 1e6, "frac_int_pipe": muls / (ms * 1e-3) / peak_mul,
                   "Mpoints_s": n / ms / 1e3, "alg_GB_s": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / hbm,
                   "note": "gpu_ms = sort + accumulate + combine + reduce per MSM; table_ms = host->device copy of the bases "
                           "+ window-table expansion, once per base set"}
            if ref is not None:
                rec.update(cpu_ms=cpu_ms, cpu_threads=threads, matches_cpu_port=(ref == out))
            print(json.dumps(rec), flush=True)
            api.msm_var(bases, sc, n, g2=g2)
            best = None
            for _ in range(2):
                o, t = api.msm_var(bases, sc, n, g2=g2)
                if best is None or t["wall_ms"] < best[1]["wall_ms"]:
                    best = (o, t)
            o, t = best
            rec = {"op": "msm_g2" if g2 else "msm_g1", "mode": "variable-base (no table)", "log_n": lg, "gpu_ms": t["kernel_ms"],
                   "wall_ms": t["wall_ms"], "GFqmul_s": muls / t["kernel_ms"] / 1e6,
                   "frac_int_pipe": muls / (t["kernel_ms"] * 1e-3) / peak_mul, "Mpoints_s": n / t["kernel_ms"] / 1e3,
                   "Mpoints_s_wall": n / t["wall_ms"] / 1e3, "same_point_as_fixed_base": (o == out),
                   "note": "one call: wall_ms = pageable host bases + scalars in, affine point out (uploads, allocation, "
                           "kernels); gpu_ms = the kernels alone; Fq-mul figure uses the canonical c = %d count" % c}
            if ref is not None:
                rec.update(cpu_ms=cpu_ms, cpu_threads=threads, matches_cpu_port=(ref == o))
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
