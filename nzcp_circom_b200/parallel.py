"""Multi-GPU plumbing: one process per GPU, torch.distributed for the (tiny) exchanges.

* Batch mode -- independent proofs sharded `i mod world`; the proving key is replicated on every GPU; NO collective on
  the data path.  The only exchange is gathering the 256-byte proofs to the caller (all_gather_object).
* Split MSM -- one large MSM partitioned by point range; each rank reduces its slice to ONE point that stays in HBM
  (nzcp_msm_plan_run_partial), a single all-gather of 128 / 256 bytes per rank (NCCL over NVLink, device buffers on
  both sides) and a kernel adds the world partials (nzcp_msm_sum_partials).  Latency-bound (~10 us), bandwidth irrelevant.

The local compute functions are parameters so that the sharding / gather logic is testable on CPU with the gloo
backend (tests/test_parallel.py); the defaults are the GPU paths of this package.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import api


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


def msm_split(bases, scalars, n_points, g2=False, group=None, device=None, mode=1, local_partial=None, sum_partials=None):
    """sum_i s_i P_i with the point range split over the ranks of `group` (BASELINE.json configs[4]).

    bases / scalars: the FULL arrays (bytes-like; every rank slices its own range).  Returns the plain affine result
    (64 / 128 bytes) on every rank.  Product path (defaults): each rank runs an MsmPlan on its slice (mode 1 =
    variable-base, no window table; mode 0 = fixed-base), which leaves ONE extended-Jacobian point (128 / 256 B) in
    HBM; one all-gather of those points over NCCL (NVLink / NVSwitch) straight from and into device memory; the world
    partials are added by a kernel (nzcp_msm_sum_partials) and only the final affine point crosses to the host.

    `local_partial(bases_slice, scalars_slice, n, g2) -> bytes` and `sum_partials(list_of_bytes, g2) -> bytes` replace the
    two GPU steps so that the slicing / gather logic runs on CPU with gloo (tests/test_parallel.py)."""
    rank, world = _world(group)
    lo, hi = shard_range(n_points, rank, world)
    bsz = 128 if g2 else 64
    bv, sv = np.frombuffer(bases, dtype=np.uint8), np.frombuffer(scalars, dtype=np.uint8)   # zero-copy views
    b_slice, s_slice = bv[lo * bsz:hi * bsz], sv[lo * 32:hi * 32]
    if (local_partial is None) != (sum_partials is None):
        raise ValueError("local_partial and sum_partials must be injected together")
    if local_partial is None:
        if device is None:
            device = rank % max(1, api.device_count())
        dev = torch.device("cuda", device)
        psz = api.XYZZ_BYTES[bool(g2)]
        mine = torch.zeros(psz, dtype=torch.uint8, device=dev)
        with api.MsmPlan(b_slice, hi - lo, g2=g2, mode=mode, device=device) as plan:
            plan.run_partial(s_slice, mine)
    else:
        mine = torch.frombuffer(bytearray(local_partial(b_slice, s_slice, hi - lo, g2)), dtype=torch.uint8)
    if world > 1:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)       # the one exchange step: one point per rank
    else:
        parts = [mine]
    if sum_partials is None:
        flat = torch.cat(parts)                          # device-to-device: world * psz contiguous bytes
        return api.msm_sum_partials(flat, world, g2=g2, device=device)
    return sum_partials([bytes(t.numpy()) for t in parts], g2)
