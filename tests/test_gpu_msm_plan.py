"""GPU parity of the round-2 MSM entry points, through the C ABI, bit-exact against the oracles:

  * nzcp_msm_var -- the variable-base (table-free) path: ffjavascript multiExpAffine's own contract, arbitrary bases per
    call; per-window bucket groups + Horner instead of the prover's window tables;
  * nzcp_msm_plan_* -- resident base sets in both modes, repeated runs, shorter scalar vectors;
  * nzcp_msm_plan_run_partial + nzcp_msm_sum_partials -- the per-rank half of the split MSM (BASELINE.json configs[4]):
    partial sums stay in HBM as XYZZ points, a kernel adds them;
  * nzcp_zkey_selfcheck -- the format facts of SURVEY.md 8c-3;
  * the pinned staging path of nzcp_prove for pageable witnesses, and the per-prover launch counter.
"""
import random

import numpy as np
import pytest

from nzcp_circom_b200 import api
from nzcp_circom_b200._lib import NzcpError
from oracle import bn254 as ob
from oracle import cref
from oracle import prover as oprover
from util import g1_plain_bytes, g2_plain_bytes, le32, tiny_case

pytestmark = pytest.mark.gpu

R = ob.R_MOD


@pytest.fixture(scope="module")
def fixed_bases():
    return ob.FixedBase(ob.G1, ob.G1_GEN), ob.FixedBase(ob.G2, ob.G2_GEN)


def _encode(g2, pts, scalars):
    enc = ob.g2_to_bytes_mont if g2 else ob.g1_to_bytes_mont
    return b"".join(enc(P) for P in pts), b"".join(le32(s) for s in scalars)


def _expected(g2, pts, scalars):
    curve = ob.G2 if g2 else ob.G1
    exp = curve.to_affine(oprover.multiexp(curve, pts, scalars))
    return g2_plain_bytes(exp) if g2 else g1_plain_bytes(exp)


def _var_case(g2, pts, scalars, window_bits=0):
    bases, sc = _encode(g2, pts, scalars)
    got, ms = api.msm_var(bases, sc, len(pts), g2=g2, window_bits=window_bits)
    assert got == _expected(g2, pts, scalars)
    assert ms["wall_ms"] >= ms["kernel_ms"] >= 0


@pytest.mark.parametrize("g2", [False, True])
@pytest.mark.parametrize("n", [0, 1, 2, 17, 300])
def test_msm_var_vs_oracle_uniform(lib, fixed_bases, g2, n):
    rng = random.Random(1000 + n * 2 + g2)
    pts = [fixed_bases[g2].mul(rng.randrange(1, R)) for _ in range(n)]
    _var_case(g2, pts, [rng.randrange(R) for _ in range(n)])


@pytest.mark.parametrize("g2", [False, True])
def test_msm_var_edge_cases(lib, fixed_bases, g2):
    """Same edge cases as the window-table path: witness-like scalars, r-1 (top signed digit carries through every
    window), infinity bases, equal and opposite points, an all-infinity sum."""
    rng = random.Random(277 + g2)
    fb = fixed_bases[g2]
    curve = ob.G2 if g2 else ob.G1
    P = fb.mul(5)
    pts = [P, P, curve.neg(P), None, fb.mul(9), None, P, fb.mul(11)] + [fb.mul(rng.randrange(1, R)) for _ in range(120)]
    scalars = [3, 3, 3, 7, 0, 0, R - 1, 1] + [rng.choice([0, 1, 1, rng.randrange(256), rng.randrange(R)]) for _ in range(120)]
    for c in (0, 2, 5, 11):
        _var_case(g2, pts, scalars, window_bits=c)
    _var_case(g2, [P] * 64, [1] * 64)
    _var_case(g2, [P] * 64, [R - 1] * 64, window_bits=4)
    _var_case(g2, [P, curve.neg(P)], [12345, 12345])
    _var_case(g2, [None] * 10, [rng.randrange(R) for _ in range(10)])


@pytest.mark.parametrize("c", [2, 3, 7, 11, 12, 13, 16])   # <= 11: all window groups fit the shared-memory histogram
def test_msm_var_window_sizes(lib, fixed_bases, c):
    rng = random.Random(500 + c)
    pts = [fixed_bases[0].mul(rng.randrange(1, R)) for _ in range(200)]
    _var_case(False, pts, [rng.randrange(R) for _ in range(190)] + [R - 1] * 10, window_bits=c)


def test_msm_var_rejects_bad_input(lib, fixed_bases):
    pts = [fixed_bases[0].mul(k + 1) for k in range(4)]
    bases, _ = _encode(False, pts, [0] * 4)
    with pytest.raises(NzcpError) as e:
        api.msm_var(bases, le32(1) + le32(R) + le32(2) + le32(3), 4)
    assert e.value.code == -7
    with pytest.raises(NzcpError) as e:
        api.msm_var(bases, le32(1) * 4, 4, window_bits=17)      # 15 windows x 2^16 buckets > 2^19
    assert e.value.code == -1


@pytest.mark.parametrize("g2,log_n", [(False, 13), (False, 17), (True, 12)])
def test_msm_var_mid_size_vs_c_oracle(lib, g2, log_n):
    """Default (cost-model) window; witness-like scalars put a quarter of the points into bucket 1 of window 0."""
    n = 1 << log_n
    bases = bytes(api.synth_points(40 + log_n, n, g2=g2))
    rng = random.Random(log_n)
    sc = b"".join(le32(rng.choice([0, 1, 1, rng.randrange(256), rng.randrange(R), rng.randrange(R)])) for _ in range(n))
    got, _ = api.msm_var(bases, sc, n, g2=g2)
    assert got == cref.msm(bases, sc, n, g2, cref.max_threads())
    # and the window-table path agrees with it (two different bucket layouts, one group element)
    assert got == api.msm(bases, sc, n, g2=g2)[0]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("g2", [False, True])
def test_msm_plan_repeated_runs(lib, fixed_bases, mode, g2):
    rng = random.Random(31 + mode + 2 * g2)
    n = 150
    pts = [fixed_bases[g2].mul(rng.randrange(1, R)) for _ in range(n)]
    pts[7] = None
    bases, _ = _encode(g2, pts, [0] * n)
    with api.MsmPlan(bases, n, g2=g2, mode=mode) as plan:
        assert plan.build_ms >= 0
        for rep in range(3):
            sc = [rng.randrange(R) for _ in range(n)]
            got, _ = plan.run(b"".join(le32(s) for s in sc))
            assert got == _expected(g2, pts, sc), (mode, rep)
        # a shorter scalar vector uses the first k bases
        k = 41
        sc = [rng.randrange(R) for _ in range(k)]
        got, _ = plan.run(b"".join(le32(s) for s in sc), k)
        assert got == _expected(g2, pts[:k], sc)
        got, _ = plan.run(b"", 0)
        assert got == bytes(128 if g2 else 64)
        with pytest.raises(NzcpError):
            plan.run(b"".join(le32(1) for _ in range(n + 1)), n + 1)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("g2", [False, True])
def test_split_msm_partials_stay_on_device(lib, fixed_bases, mode, g2):
    """Three 'ranks' on one GPU: each slice's partial is written as an XYZZ point into its slot of a device buffer (what
    NCCL all-gathers), the sum kernel adds them -> equals the one-piece MSM and the oracle."""
    import torch
    rng = random.Random(91 + mode + 2 * g2)
    n, world = 333, 3
    pts = [fixed_bases[g2].mul(rng.randrange(1, R)) for _ in range(n)]
    sc = [rng.randrange(R) for _ in range(n)]
    psz = api.XYZZ_BYTES[g2]
    bsz = 128 if g2 else 64
    bases, scb = _encode(g2, pts, sc)
    buf = torch.zeros(world * psz, dtype=torch.uint8, device="cuda:0")
    from nzcp_circom_b200 import parallel
    for rank in range(world):
        lo, hi = parallel.shard_range(n, rank, world)
        with api.MsmPlan(bases[lo * bsz:hi * bsz], hi - lo, g2=g2, mode=mode) as plan:
            plan.run_partial(scb[lo * 32:hi * 32], buf.data_ptr() + rank * psz)
    torch.cuda.synchronize()
    got = api.msm_sum_partials(buf, world, g2=g2)
    assert got == _expected(g2, pts, sc)
    # world size 1 through the public helper (no process group): the same product path end to end
    assert parallel.msm_split(bases, scb, n, g2=g2, mode=mode) == got
    # an empty slice contributes the point at infinity
    with api.MsmPlan(b"", 0, g2=g2, mode=mode) as plan:
        plan.run_partial(b"", buf.data_ptr(), 0)
    lo, hi = parallel.shard_range(n, 0, world)
    assert api.msm_sum_partials(buf, world, g2=g2) == _expected(g2, pts[hi:], sc[hi:])


def test_split_msm_large_vs_c_oracle(lib):
    """2^18 dense points in 4 slices, variable-base plans, device-side sum -- against the C oracle."""
    import torch
    from nzcp_circom_b200 import parallel
    n, world = 1 << 18, 4
    bases = bytes(api.synth_points(555, n))
    rs = np.random.RandomState(9)
    sc = rs.randint(0, 2 ** 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    sc[:, 7] &= 0x1FFFFFFF
    scb = sc.tobytes()
    buf = torch.zeros(world * 128, dtype=torch.uint8, device="cuda:0")
    for rank in range(world):
        lo, hi = parallel.shard_range(n, rank, world)
        with api.MsmPlan(bases[lo * 64:hi * 64], hi - lo, mode=1) as plan:
            plan.run_partial(scb[lo * 32:hi * 32], buf.data_ptr() + rank * 128)
    assert api.msm_sum_partials(buf, world) == cref.msm(bases, scb, n, False, cref.max_threads())


def test_zkey_selfcheck(lib):
    c = tiny_case(seed=5, n_constraints=300, n_public=7, n_free=20)
    zb = bytearray(c["zkey_bytes"])
    rep = api.zkey_selfcheck(zb)
    assert rep["ok"], rep
    assert rep["n_public"] == 7 and rep["n_vars"] == c["n_vars"] and rep["n_constraints"] == 300
    assert rep["public_rows_ok"] and rep["header_points_ok"] and rep["header_moduli_ok"]
    assert sum(rep["off_curve"].values()) == 0 and rep["bad_coef_values"] == 0
    assert rep["infinity"]["H"] == 0
    from oracle import formats
    secs = {sid: v[0] for sid, v in formats.read_container(zb, b"zkey")[1].items()}
    # one flipped bit in a base point of section 5 (A) and of section 7 (B2)
    for sec, name in ((5, "A"), (7, "B2"), (9, "H")):
        bad = bytearray(zb)
        off, ln = secs[sec]
        # first non-infinity point of the section
        psz = 128 if sec == 7 else 64
        k = next(i for i in range(ln // psz) if any(bad[off + i * psz: off + (i + 1) * psz]))
        bad[off + k * psz + 3] ^= 0x10
        rep = api.zkey_selfcheck(bad)
        assert not rep["ok"] and rep["off_curve"][name] == 1, (name, rep)
    # the appended public-input rows: break the last coefficient value (R^2 mod r)
    bad = bytearray(zb)
    off, ln = secs[4]
    bad[off + ln - 32] ^= 1
    rep = api.zkey_selfcheck(bad)
    assert not rep["ok"] and not rep["public_rows_ok"]
    # a coefficient >= r
    bad = bytearray(zb)
    bad[off + 4 + 12: off + 4 + 44] = le32(R)
    rep = api.zkey_selfcheck(bad)
    assert not rep["ok"] and rep["bad_coef_values"] == 1
    # not a zkey at all
    with pytest.raises(NzcpError):
        api.zkey_selfcheck(b"wtns" + bytes(60))


def test_selfcheck_on_benchmarked_key(lib):
    """The GPU-made synthetic key of the nzcp_exampleTest shape passes every format check (0.5 GB of base points)."""
    from util import TOXIC
    sc = api.SynthCircuit(seed=0xC0FFEE, n_constraints=716000, n_public=513, n_free=44000)
    zb = sc.zkey([TOXIC[k] for k in ("tau", "alpha", "beta", "gamma", "delta")])
    rep = api.zkey_selfcheck(zb)
    assert rep["ok"], rep
    assert rep["n_constraints"] == 716000 and rep["domain_size"] == 1 << 20 and rep["n_public"] == 513


def test_pageable_and_pinned_witness_paths_agree(lib):
    """nzcp_prove from pageable memory goes through the prover's pinned staging buffer (chunked), from pinned memory
    straight to the copy engine; forced modes 0 / 1 and tiny chunks give the same proof."""
    import torch
    c = tiny_case(seed=21, n_constraints=2500, n_public=9, n_free=40)
    r, s = 0x1234567, 0x7654321
    wt = bytes(c["wtns_bytes"])
    pinned = torch.frombuffer(bytearray(wt), dtype=torch.uint8).pin_memory()
    with api.Zkey(c["zkey_bytes"]) as zk, api.Prover(zk) as pr:
        base = pr.prove(wt, r=r, s=s)["proof"]
        try:
            for mode, chunk in ((0, 1024), (1, 1024), (1, 64), (-1, 64)):
                api.tuning_set("stage_mode", mode)
                api.tuning_set("stage_chunk_kb", chunk)
                assert pr.prove(wt, r=r, s=s)["proof"] == base, (mode, chunk)
                assert pr.prove(pinned.numpy(), r=r, s=s)["proof"] == base, (mode, chunk, "pinned")
        finally:
            api.tuning_set("stage_mode", -1)
            api.tuning_set("stage_chunk_kb", 1024)
    exp, _ = oprover.prove_files(c["zkey_bytes"], wt, r, s)
    assert base == oprover.proof_to_bytes(exp)


def test_launch_count_is_per_prover(lib):
    c = tiny_case(seed=22, n_constraints=400, n_public=3, n_free=10)
    with api.Zkey(c["zkey_bytes"]) as zk, api.Prover(zk) as p1, api.Prover(zk) as p2:
        assert p1.launch_count() == 0 and p2.launch_count() == 0
        p1.prove(c["wtns_bytes"], r=1, s=2)
        n1 = p1.launch_count()
        assert n1 > 20 and p2.launch_count() == 0
        p2.prove(c["wtns_bytes"], r=1, s=2)
        p2.prove(c["wtns_bytes"], r=1, s=2)
        assert p1.launch_count() == n1 and p2.launch_count() == 2 * n1
