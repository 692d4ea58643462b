"""GPU parity at the sizes bench.py REPORTS (BASELINE.json configs[1], [2], [3]) -- not through size-independent
properties but value by value against the C restatement of snarkjs (oracle/c, multi-threaded so it finishes in seconds):

  * one proof of the nzcp_liveTest shape (the benchmarked workload) and one of the nzcp_exampleTest shape:
    the H "coefficients" (joinABC output), each of the five MSM results and the 256-byte proof, bit for bit, r/s fixed
    (north_star: "H coefficients, each MSM result and the final proof");
  * Fr.fft / Fr.ifft and the coset pipeline (batch 3, as the prover runs it) at 2^20 and 2^22;
  * one dense 2^20 G1 MSM and one dense 2^18 G2 MSM.

The C oracle itself is pinned to the Python restatement and the toxic-waste closed form in the CPU suite
(tests/test_oracle.py, tests/test_golden.py); "parity unpinned" vs snarkjs itself (no node here) applies as everywhere.
"""
import numpy as np
import pytest

from nzcp_circom_b200 import api, groth16
from oracle import cref
from util import TOXIC

pytestmark = pytest.mark.gpu

SHAPES = {
    # bench.py SHAPES: (n_constraints, n_public, n_free)
    "live": (835000, 513, 45000),
    "example": (716000, 513, 44000),
}
R_FIXED = 0x0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF0123456789ABCD
S_FIXED = 0x0FEDCBA9876543210FEDCBA9876543210FEDCBA9876543210FEDCBA9876543


@pytest.mark.parametrize("shape", ["live", "example"])
def test_benchmarked_shape_h_msms_proof_bit_exact(lib, shape):
    nc, npub, nfree = SHAPES[shape]
    sc = api.SynthCircuit(seed=0xC0FFEE, n_constraints=nc, n_public=npub, n_free=nfree)     # bench.py make_workload
    zb = sc.zkey([TOXIC[k] for k in ("tau", "alpha", "beta", "gamma", "delta")])
    n = sc.domain_size
    assert n == 1 << 20 and sc.n_public == 513
    with api.Zkey(zb) as zk, api.Prover(zk) as pr:
        for seed in (1, 2):
            wb = sc.wtns(seed)
            got = pr.prove(wb, r=R_FIXED, s=S_FIXED, debug=True, want_h=True)
            exp = cref.prove(zb, wb, R_FIXED, S_FIXED, want_h_size=n)
            assert got["h"] == exp["h"], "H scalars (joinABC output) differ"
            for k in ("msm_a", "msm_b1", "msm_b2", "msm_c", "msm_h"):
                assert got[k] == exp[k], k
            assert got["proof"] == exp["proof"]
            if seed == 1:
                first = got["proof"]
        assert got["proof"] != first
        # throughput-mode prover (three batched-affine pair rounds on every MSM at this size): the same five MSM results
        with api.Prover(zk, throughput=True) as pt:
            got_t = pt.prove(wb, r=R_FIXED, s=S_FIXED, debug=True)
            for k in ("msm_a", "msm_b1", "msm_b2", "msm_c", "msm_h", "proof"):
                assert got_t[k] == exp[k], "throughput mode: " + k
        # the same witness through the batch entry point (several provers in flight) gives the same bytes
        assert zk.prove_batch([wb, wb, wb], [R_FIXED] * 3, [S_FIXED] * 3, n_provers=3) == [got["proof"]] * 3
    # and it is a valid Groth16 proof (independent pairing check)
    out = groth16.prove({"type": "mem", "data": zb}, {"type": "mem", "data": wb}, r=R_FIXED, s=S_FIXED)
    assert groth16.verify(groth16.exportVerificationKey(zb), out["publicSignals"], out["proof"])
    groth16.terminate()


def _rand_fr_mont(n, seed):
    """n x 8 uint32 limbs, each value < 2^253 < r (canonical).  Any canonical value is a valid Montgomery-form input."""
    rs = np.random.RandomState(seed)
    x = rs.randint(0, 2 ** 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    x[:, 7] &= 0x1FFFFFFF
    return x


@pytest.mark.parametrize("log_n", [20, 22])
def test_ntt_full_size_vs_c_oracle(lib, log_n):
    n = 1 << log_n
    thr = cref.max_threads()
    x = _rand_fr_mont(n, log_n)
    for inverse in (False, True):
        g, e = x.copy(), x.copy()
        api.ntt(g, log_n, inverse=inverse)
        cref.ntt(e, log_n, inverse=inverse, threads=thr)
        assert np.array_equal(g, e), "ntt inverse=%s 2^%d" % (inverse, log_n)


@pytest.mark.parametrize("log_n", [20, 22])
def test_ntt_coset_batch3_full_size_vs_c_oracle(lib, log_n):
    n = 1 << log_n
    thr = cref.max_threads()
    x = _rand_fr_mont(3 * n, 100 + log_n)
    g = x.copy()
    api.ntt_coset(g, log_n, batch=3)
    for k in range(3):
        e = np.ascontiguousarray(x[k * n:(k + 1) * n])
        cref.ntt_coset(e, log_n, threads=thr)
        assert np.array_equal(g[k * n:(k + 1) * n], e), "polynomial %d" % k


@pytest.mark.parametrize("g2,log_n", [(False, 20), (True, 18)])
def test_msm_dense_full_size_vs_c_oracle(lib, g2, log_n):
    n = 1 << log_n
    bases = bytes(api.synth_points(31 + log_n, n, g2=g2))
    sc = _rand_fr_mont(n, 7 + log_n)                       # uniform 253-bit plain scalars
    sc[0] = 0
    sc[1] = 0
    sc[1, 0] = 1
    got, _ = api.msm(bases, sc, n, g2=g2)
    assert got == cref.msm(bases, sc, n, g2, cref.max_threads())


@pytest.mark.parametrize("rounds", [3])
def test_msm_dense_full_size_with_pair_rounds(lib, rounds):
    """The opt-in batched-affine pair rounds at 2^20 dense points (8.4 M + 4.2 M + 2.1 M affine additions sharing
    inversions, then the XYZZ tail) against the C oracle."""
    n = 1 << 20
    bases = bytes(api.synth_points(77, n))
    sc = _rand_fr_mont(n, 1234)
    try:
        api.tuning_set("msm_rounds", rounds)
        got, _ = api.msm(bases, sc, n)
    finally:
        api.tuning_set("msm_rounds", -1)
    assert got == cref.msm(bases, sc, n, False, cref.max_threads())
