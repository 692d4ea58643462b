"""CPU, world_size 2, gloo: the N>1 host logic -- batch sharding with proof gather, and the split MSM with its one
all-gather.  Local compute is injected (the C oracle stands in for the GPU kernels, as a test double only)."""
import os
import random
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from nzcp_circom_b200 import parallel
from oracle import bn254 as ob
from oracle import cref
from util import g1_plain_bytes, g2_plain_bytes, le32

R = ob.R_MOD


def test_shard_helpers():
    for n in (0, 1, 7, 8, 1024):
        for world in (1, 2, 3, 8):
            idx = sorted(i for r in range(world) for i in parallel.shard_indices(n, r, world))
            assert idx == list(range(n))
            ranges = [parallel.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def _sum_affine(parts, g2):
    """Test double of nzcp_msm_sum_partials: the injected partials are plain affine points."""
    curve = ob.G2 if g2 else ob.G1
    dec = (lambda b: None if not any(b) else ((int.from_bytes(b[0:32], "little"), int.from_bytes(b[32:64], "little")),
                                              (int.from_bytes(b[64:96], "little"), int.from_bytes(b[96:128], "little")))) \
        if g2 else (lambda b: None if not any(b) else (int.from_bytes(b[0:32], "little"), int.from_bytes(b[32:64], "little")))
    acc = None
    for p in parts:
        acc = curve.add(acc, dec(p))
    return g2_plain_bytes(acc) if g2 else g1_plain_bytes(acc)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- batch mode: 7 "witnesses", fake prover tags each proof with its index and the rank that made it
        wt = [b"w%d" % i for i in range(7)]
        seen = []

        def prove_many(indices):
            seen.extend(indices)
            return [bytes([i, rank]) * 128 for i in indices]
        proofs = parallel.prove_batch(None, wt, prove_many=prove_many)
        assert seen == list(range(rank, 7, world))
        assert [p[0] for p in proofs] == list(range(7))
        assert [p[1] for p in proofs] == [i % world for i in range(7)]
        # ---- split MSM, G1 and G2, odd point count (uneven slices)
        for g2 in (False, True):
            rng = random.Random(5 + g2)
            curve = ob.G2 if g2 else ob.G1
            fb = ob.FixedBase(curve, curve.gen)
            n = 37
            pts = [fb.mul(rng.randrange(1, R)) for _ in range(n)]
            sc = [rng.randrange(R) for _ in range(n)]
            enc = ob.g2_to_bytes_mont if g2 else ob.g1_to_bytes_mont
            bases = b"".join(enc(p) for p in pts)
            scalars = b"".join(le32(s) for s in sc)
            got = parallel.msm_split(bases, scalars, n, g2=g2,
                                     local_partial=lambda b, s, k, g: cref.msm(bytes(b), bytes(s), k, g, 1),
                                     sum_partials=_sum_affine)
            full = cref.msm(bases, scalars, n, g2, 1)
            assert got == full
            naive = None
            for P, k in zip(pts, sc):
                naive = curve.add(naive, curve.mul(P, k))
            assert got == (g2_plain_bytes(naive) if g2 else g1_plain_bytes(naive))
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL: %r" % (e,)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
