"""Multi-GPU plumbing: one process per GPU, torch.distributed for the (tiny) exchanges.

* Batch mode -- independent proofs sharded `i mod world`; the proving key is replicated on every GPU; NO collective on
  the data path.  The only exchange is gathering the 256-byte proofs to the caller (all_gather_object).
* Split MSM -- one large MSM partitioned by point range; each rank reduces its slice to ONE point with the full
  single-GPU pipeline (nzcp_msm), then a single all-gather of 64 / 128 bytes per rank (NCCL over NVLink when the
  backend is nccl) and world-1 group additions on every rank.  Latency-bound (~10 us), bandwidth irrelevant.

The local compute functions are parameters so that the sharding / gather logic is testable on CPU with the gloo
backend (tests/test_parallel.py); the defaults are the GPU paths of this package.
"""
import torch
import torch.distributed as dist

from . import api, verifier


def shard_indices(n_items, rank, world):
    """Batch mode: proof i goes to rank i mod world."""
    return list(range(rank, n_items, world))


def shard_range(n_points, rank, world):
    """Split MSM: contiguous point slice of this rank (sizes differ by at most one)."""
    base, rem = divmod(n_points, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def prove_batch(zkey, wtns_list, r_list=None, s_list=None, device=None, n_provers=2, group=None, prove_many=None):
    """Prove every witness of `wtns_list` (same list on all ranks) across the ranks of `group`.

    Returns the full list of 256-byte proofs on every rank, in input order.  `prove_many(indices) -> list of proof
    bytes` may be supplied (tests); the default keeps a ProverPool on this rank's GPU."""
    rank, world = _world(group)
    mine = shard_indices(len(wtns_list), rank, world)
    if prove_many is None:
        if device is None:
            device = rank % max(1, api.device_count())
        zk = zkey if isinstance(zkey, api.Zkey) else api.Zkey(zkey, device)
        pool = api.ProverPool(zk, n_provers)
        try:
            # per-proof r, s (None = random, as snarkjs) -- run in pool-sized waves to keep per-proof scalars
            local = []
            k = len(pool.provers)
            for w0 in range(0, len(mine), k):
                chunk = mine[w0:w0 + k]
                futs = [pool._pool.submit(pool.provers[j].prove, wtns_list[i],
                                          None if r_list is None else r_list[i],
                                          None if s_list is None else s_list[i])
                        for j, i in enumerate(chunk)]
                local += [f.result()["proof"] for f in futs]
        finally:
            pool.close()
            if not isinstance(zkey, api.Zkey):
                zk.close()
    else:
        local = prove_many(mine)
    if world == 1:
        return local
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, local), group=group)
    out = [None] * len(wtns_list)
    for idx, proofs in gathered:
        for i, p in zip(idx, proofs):
            out[i] = p
    return out


def _point_from_bytes(b, g2):
    le = lambda o: int.from_bytes(b[o:o + 32], "little")  # noqa: E731
    if g2:
        v = [le(32 * i) for i in range(4)]
        return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))
    x, y = le(0), le(32)
    return None if x == 0 and y == 0 else (x, y)


def _point_to_bytes(P, g2):
    if P is None:
        return bytes(128 if g2 else 64)
    flat = [P[0][0], P[0][1], P[1][0], P[1][1]] if g2 else [P[0], P[1]]
    return b"".join(int(v).to_bytes(32, "little") for v in flat)


def msm_split(bases, scalars, n_points, g2=False, group=None, device=None, local_msm=None):
    """sum_i s_i P_i with the point range split over the ranks of `group`.

    bases / scalars: the FULL arrays (bytes-like; every rank slices its own range).  Returns the plain affine result
    (64 / 128 bytes) on every rank.  `local_msm(bases_slice, scalars_slice, n, g2) -> point bytes` defaults to the
    single-GPU nzcp_msm."""
    rank, world = _world(group)
    lo, hi = shard_range(n_points, rank, world)
    bsz = 128 if g2 else 64
    bv, sv = memoryview(bases).cast("B"), memoryview(scalars).cast("B")
    if local_msm is None:
        if device is None:
            device = rank % max(1, api.device_count())

        def local_msm(b, s, n, g2_):
            return api.msm(bytes(b), bytes(s), n, g2=g2_, device=device)[0]
    part = local_msm(bv[lo * bsz:hi * bsz], sv[lo * 32:hi * 32], hi - lo, g2)
    if world == 1:
        return bytes(part)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(part), dtype=torch.uint8).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)       # the one exchange step: bsz bytes per rank
    add = verifier.g2_add if g2 else verifier.g1_add
    acc = None
    for t in parts:
        acc = add(acc, _point_from_bytes(bytes(t.cpu().numpy()), g2))
    return _point_to_bytes(acc, g2)
