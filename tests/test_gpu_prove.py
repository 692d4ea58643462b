"""GPU parity for the whole path: groth16.prove through the C ABI vs the oracle's restatement of snarkjs
groth16_prove.js -- H scalars, each of the five MSM results and the final proof, bit for bit, with r and s fixed.

Also the error behaviour snarkjs has at this boundary, and full-size (n = 2^20) property checks: the proof must satisfy
the pairing equation and pi_a / pi_b must equal the toxic-waste closed form (which pins pi_c too: given pi_a, pi_b and
the public signals the verification equation has exactly one solution for pi_c).
"""
import random
import struct

import pytest

from nzcp_circom_b200 import api, groth16
from nzcp_circom_b200._lib import NzcpError
from oracle import bn254 as ob
from oracle import formats, setup
from oracle import prover as oprover
from util import TOXIC, g1_plain_bytes, g2_plain_bytes, le32, tiny_case

pytestmark = pytest.mark.gpu
R = ob.R_MOD

CASES = [
    # seed, n_constraints, n_public, n_free       -> domain size
    (1, 1, 1, 16),        # 4 (smallest the format allows: 1 constraint + 2 public rows)
    (2, 5, 2, 16),        # 8
    (3, 28, 3, 16),       # 32  exactly full: 28 + 3 + 1
    (4, 29, 3, 16),       # 64  one over
    (5, 200, 5, 20),      # 256
    (6, 700, 13, 30),     # 1024
    (7, 1500, 0, 16),     # 2048, no public signals
]


@pytest.mark.parametrize("seed,nc,npub,nfree", CASES)
def test_prove_matches_oracle_bit_exact(lib, seed, nc, npub, nfree):
    c = tiny_case(seed, nc, npub, nfree)
    rng = random.Random(seed)
    r, s = rng.randrange(R), rng.randrange(R)
    zkd = formats.read_zkey(c["zkey_bytes"])
    exp, pub, parts = oprover.prove(zkd, c["witness"], r, s, return_parts=True)
    with api.Zkey(c["zkey_bytes"]) as zk, api.Prover(zk) as pr:
        assert (zk.n_vars, zk.n_public, zk.domain_size) == (c["n_vars"], npub, zkd["domainSize"])
        got = pr.prove(c["wtns_bytes"], r=r, s=s, debug=True, want_h=True)
        # H "coefficients" (joinABC output = the section-9 MSM scalars)
        assert got["h"] == b"".join(le32(x) for x in parts["h"])
        assert got["msm_a"] == g1_plain_bytes(parts["A"])
        assert got["msm_b1"] == g1_plain_bytes(parts["B1"])
        assert got["msm_b2"] == g2_plain_bytes(parts["B2"])
        assert got["msm_c"] == g1_plain_bytes(parts["C"])
        assert got["msm_h"] == g1_plain_bytes(parts["H"])
        assert got["proof"] == oprover.proof_to_bytes(exp)
        # same proof again from the bare-witness entry point, and a second run is deterministic
        body = b"".join(le32(w) for w in c["witness"])
        assert pr.prove_witness(body, c["n_vars"], r=r, s=s)["proof"] == got["proof"]
    # closed form from the toxic waste: independent of the NTT/MSM path in both implementations
    cf = setup.expected_proof(c["constraints"], c["n_vars"], npub, c["toxic"], c["witness"], r, s)
    assert got["proof"] == oprover.proof_to_bytes(cf)


def test_snarkjs_surface_and_verify(lib):
    c = tiny_case(21, 120, 4, 16)
    zkey = {"type": "mem", "data": c["zkey_bytes"]}
    out = groth16.prove(zkey, {"type": "mem", "data": c["wtns_bytes"]}, r=5, s=7)
    exp, pub = oprover.prove(formats.read_zkey(c["zkey_bytes"]), c["witness"], 5, 7)
    assert out["proof"] == oprover.proof_to_json(exp)
    assert out["publicSignals"] == [str(x) for x in pub] == [str(x) for x in c["witness"][1:5]]
    vk = groth16.exportVerificationKey(c["zkey_bytes"])
    assert groth16.verify(vk, out["publicSignals"], out["proof"])
    bad = list(out["publicSignals"])
    bad[0] = str((int(bad[0]) + 1) % R)
    assert not groth16.verify(vk, bad, out["proof"])
    # random blinding (r = s = None -> OS CSPRNG, as Fr.random()): different proofs, both verify
    o1 = groth16.prove(zkey, {"type": "mem", "data": c["wtns_bytes"]})
    o2 = groth16.prove(zkey, {"type": "mem", "data": c["wtns_bytes"]})
    assert o1["proof"] != o2["proof"]
    assert groth16.verify(vk, o1["publicSignals"], o1["proof"]) and groth16.verify(vk, o2["publicSignals"], o2["proof"])
    # fullProve with a witness-calculator callable
    o3 = groth16.fullProve({"x": 1}, lambda inp: c["wtns_bytes"], zkey, r=5, s=7)
    assert o3 == out
    groth16.terminate()


def test_prove_unsatisfying_witness_fails_verification(lib):
    c = tiny_case(22, 60, 2, 16)
    w = list(c["witness"])
    w[-1] = (w[-1] + 1) % R
    out = groth16.prove({"type": "mem", "data": c["zkey_bytes"]}, {"type": "mem", "data": formats.write_wtns(w)}, r=1, s=2)
    # still bit-exact with the oracle (snarkjs proves garbage without complaint) ...
    exp, _ = oprover.prove(formats.read_zkey(c["zkey_bytes"]), w, 1, 2)
    assert out["proof"] == oprover.proof_to_json(exp)
    # ... but the pairing check rejects it
    assert not groth16.verify(groth16.exportVerificationKey(c["zkey_bytes"]), out["publicSignals"], out["proof"])
    groth16.terminate()


def test_errors_match_snarkjs(lib):
    c = tiny_case(23, 10, 1, 16)
    zb, wb = bytearray(c["zkey_bytes"]), bytearray(c["wtns_bytes"])
    with api.Zkey(zb) as zk, api.Prover(zk) as pr:
        # witness length
        short = formats.write_wtns(c["witness"][:-1])
        with pytest.raises(NzcpError, match=r"Invalid witness length. Circuit: %d, witness: %d" % (c["n_vars"], c["n_vars"] - 1)) as e:
            pr.prove(short)
        assert e.value.code == -5
        # witness over another field
        w2 = bytearray(wb)
        w2[12 + 12 + 4] ^= 1
        with pytest.raises(NzcpError, match="Curve of the witness does not match the curve of the proving key") as e:
            pr.prove(w2)
        assert e.value.code == -4
        # non-canonical witness value
        w3 = bytearray(wb)
        off = len(w3) - 32
        w3[off:off + 32] = le32(R)
        with pytest.raises(NzcpError) as e:
            pr.prove(w3, r=1, s=1)
        assert e.value.code == -7
        # bad magic / truncated
        with pytest.raises(NzcpError) as e:
            pr.prove(b"wtnx" + bytes(wb[4:]))
        assert e.value.code == -2
        with pytest.raises(NzcpError) as e:
            pr.prove(bytes(wb[:100]))
        assert e.value.code == -2
        # r >= field order
        with pytest.raises(NzcpError) as e:
            pr.prove(wb, r=R, s=1)
        assert e.value.code == -1
        # the prover is still usable after errors
        assert len(pr.prove(wb, r=1, s=1)["proof"]) == 256
    # zkey that is not groth16 (protocol id 2 = plonk)
    z2 = bytearray(zb)
    z2[12 + 12:12 + 12 + 4] = struct.pack("<I", 2)
    with pytest.raises(NzcpError, match="zkey file is not groth16") as e:
        api.Zkey(z2)
    assert e.value.code == -3
    with pytest.raises(NzcpError) as e:
        api.Zkey(b"zkex" + bytes(zb[4:]))
    assert e.value.code == -2
    with pytest.raises(NzcpError) as e:
        api.Zkey(bytes(zb[:len(zb) // 2]))
    assert e.value.code == -2


def test_gpu_synthetic_setup_matches_oracle_setup(lib):
    """The CUDA fixed-base setup (synth.cu) and the Python setup give byte-identical .zkey files for the same R1CS and
    toxic waste; the synthetic witness satisfies the R1CS."""
    sc = api.SynthCircuit(seed=3, n_constraints=120, n_public=6, n_free=24)
    toxic = [TOXIC[k] for k in ("tau", "alpha", "beta", "gamma", "delta")]
    zb = sc.zkey(toxic)
    r1 = formats.read_r1cs(bytes(sc.r1cs()))
    assert r1["nConstraints"] == 120 and r1["nPublic"] == 6 and r1["nVars"] == sc.n_vars
    zk_py = setup.make_zkey(r1["constraints"], r1["nVars"], 6, TOXIC)
    a, b = formats.read_zkey(bytes(zb)), formats.read_zkey(formats.write_zkey(zk_py))
    for k in ("nVars", "nPublic", "domainSize", "vk_alpha_1", "vk_beta_1", "vk_beta_2", "vk_gamma_2", "vk_delta_1",
              "vk_delta_2", "IC", "A", "B1", "B2", "C", "H"):
        assert a[k] == b[k], k
    assert sorted(a["coefs"]) == sorted(b["coefs"])
    wt = formats.read_wtns(bytes(sc.wtns(5)))["witness"]
    for (A, B, C) in r1["constraints"]:
        ev = [sum(v * wt[s] for s, v in lc.items()) % R for lc in (A, B, C)]
        assert ev[0] * ev[1] % R == ev[2]
    # and the proof from those files matches oracle + closed form
    out = groth16.prove({"type": "mem", "data": zb}, {"type": "mem", "data": sc.wtns(5)}, r=11, s=13)
    exp, _ = oprover.prove(a, wt, 11, 13)
    assert out["proof"] == oprover.proof_to_json(exp)
    groth16.terminate()


def _closed_form_ab(r1cs_bytes, n_public, toxic, witness_bytes, r, s):
    """pi_a, pi_b from the toxic waste, streaming the .r1cs (numpy-free, O(nnz) big-int work)."""
    tau, alpha, beta, delta = (toxic[k] % R for k in ("tau", "alpha", "beta", "delta"))
    r1 = formats.read_r1cs(r1cs_bytes)
    nc = r1["nConstraints"]
    lg = setup.domain_log(nc, n_public)
    n = 1 << lg
    # Lagrange basis at tau with one batched inversion
    w = ob.FR_W[lg]
    zt = (pow(tau, n, R) - 1) * pow(n, -1, R) % R
    wj, den, ws = 1, [], []
    for _ in range(nc + n_public + 1):
        ws.append(wj)
        den.append((tau - wj) % R)
        wj = wj * w % R
    pre, acc = [], 1
    for d in den:
        pre.append(acc)
        acc = acc * d % R
    inv = pow(acc, -1, R)
    L = [0] * len(den)
    for i in range(len(den) - 1, -1, -1):
        L[i] = zt * ws[i] % R * (inv * pre[i] % R) % R
        inv = inv * den[i] % R
    wt = formats.read_wtns(witness_bytes)["witness"]
    a = b = 0
    for j, (A, B, _C) in enumerate(r1["constraints"]):
        a += L[j] * sum(v * wt[sg] for sg, v in A.items())
        b += L[j] * sum(v * wt[sg] for sg, v in B.items())
    for i in range(n_public + 1):
        a += L[nc + i] * wt[i]
    fb1 = ob.FixedBase(ob.G1, ob.G1_GEN)
    fb2 = ob.FixedBase(ob.G2, ob.G2_GEN)
    return fb1.mul((alpha + a + r * delta) % R), fb2.mul((beta + b + s * delta) % R)


@pytest.mark.parametrize("nc,npub,nfree", [(20000, 33, 500), (716000, 513, 43487)])
def test_full_size_proof_verifies_and_matches_closed_form(lib, nc, npub, nfree):
    """Size-independent properties at n = 2^15 and at the nzcp_exampleTest shape (n = 2^20, m = 760 001, nPublic = 513):
    pairing equation holds; pi_a, pi_b equal the closed form => pi_c is the unique solution => the proof is exact."""
    sc = api.SynthCircuit(seed=42, n_constraints=nc, n_public=npub, n_free=nfree)
    toxic = [TOXIC[k] for k in ("tau", "alpha", "beta", "gamma", "delta")]
    zb = sc.zkey(toxic)
    wb = sc.wtns(1)
    r, s = 0xABCDEF0123456789 << 100, 0x13579BDF02468ACE << 90
    zsrc = {"type": "mem", "data": zb}
    out = groth16.prove(zsrc, {"type": "mem", "data": wb}, r=r, s=s)
    vk = groth16.exportVerificationKey(zb)
    assert len(out["publicSignals"]) == npub
    assert groth16.verify(vk, out["publicSignals"], out["proof"])
    pa, pb = _closed_form_ab(bytes(sc.r1cs()), npub, TOXIC, bytes(wb), r, s)
    assert out["proof"]["pi_a"][:2] == [str(pa[0]), str(pa[1])]
    assert out["proof"]["pi_b"][:2] == [[str(pb[0][0]), str(pb[0][1])], [str(pb[1][0]), str(pb[1][1])]]
    # a second witness through the same resident key
    out2 = groth16.prove(zsrc, {"type": "mem", "data": sc.wtns(2)}, r=r, s=s)
    assert out2["proof"] != out["proof"]
    assert groth16.verify(vk, out2["publicSignals"], out2["proof"])
    groth16.terminate()


def test_prover_pool_matches_single_prover(lib):
    """Several proofs in flight on one GPU (host threads + per-prover streams) give the same bytes as one at a time."""
    sc = api.SynthCircuit(seed=5, n_constraints=3000, n_public=9, n_free=64)
    zb = sc.zkey([TOXIC[k] for k in ("tau", "alpha", "beta", "gamma", "delta")])
    wl = [sc.wtns(100 + k) for k in range(7)]
    with api.Zkey(zb) as zk:
        with api.Prover(zk) as pr:
            single = [pr.prove(w, r=3 + k, s=5 + k)["proof"] for k, w in enumerate(wl)]
        with api.ProverPool(zk, 3) as pool:
            many = [pool.provers[k % 3].prove(w, r=3 + k, s=5 + k)["proof"] for k, w in enumerate(wl)]
            same_rs = pool.prove_many(wl, r=11, s=13)
        assert many == single
        with api.Prover(zk) as pr:
            assert [x["proof"] for x in same_rs] == [pr.prove(w, r=11, s=13)["proof"] for w in wl]
    vk = groth16.exportVerificationKey(zb)
    out = groth16.prove({"type": "mem", "data": zb}, {"type": "mem", "data": wl[0]}, r=11, s=13)
    assert groth16.verify(vk, out["publicSignals"], out["proof"])
    groth16.terminate()


def test_prove_batch_c_abi(lib):
    """nzcp_prove_batch (library-side thread pool) == one nzcp_prove per witness; per-proof r, s; errors are reported."""
    sc = api.SynthCircuit(seed=6, n_constraints=2000, n_public=7, n_free=50)
    zb = sc.zkey([TOXIC[k] for k in ("tau", "alpha", "beta", "gamma", "delta")])
    wl = [sc.wtns(200 + k) for k in range(8)]
    rs = [(17 + k, 29 + 3 * k) for k in range(8)]
    with api.Zkey(zb) as zk:
        with api.Prover(zk) as pr:
            single = [pr.prove(w, r=r, s=s)["proof"] for w, (r, s) in zip(wl, rs)]
        for n_prov in (1, 3):
            assert zk.prove_batch(wl, [r for r, _ in rs], [s for _, s in rs], n_provers=n_prov) == single
        assert zk.prove_batch([], None, None) == []
        rnd = zk.prove_batch(wl[:2])                       # random blinding
        assert rnd[0] != single[0] and len(rnd[0]) == 256
        bad = list(wl)
        bad[5] = formats.write_wtns(formats.read_wtns(bytes(wl[5]))["witness"][:-1])
        with pytest.raises(NzcpError, match="proof 5: Invalid witness length") as e:
            zk.prove_batch(bad, [r for r, _ in rs], [s for _, s in rs])
        assert e.value.code == -5
        assert zk.prove_batch(wl, [r for r, _ in rs], [s for _, s in rs]) == single     # pool still healthy


def test_two_devices_in_one_process(lib):
    """The C ABI takes a device index: keys on two GPUs of one process give the same proofs (kernel attributes such as
    the > 48 KB shared-memory opt-in are per device)."""
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    c = tiny_case(41, 700, 5, 16)
    outs = []
    for dev in (0, 1):
        with api.Zkey(c["zkey_bytes"], device=dev) as zk, api.Prover(zk) as pr:
            outs.append(pr.prove(c["wtns_bytes"], r=9, s=10, debug=True, want_h=True))
    assert outs[0]["proof"] == outs[1]["proof"] and outs[0]["h"] == outs[1]["h"]
    exp, _ = oprover.prove(formats.read_zkey(c["zkey_bytes"]), c["witness"], 9, 10)
    assert outs[1]["proof"] == oprover.proof_to_bytes(exp)


@pytest.mark.parametrize("rounds", [1, 3])
def test_prove_with_pair_rounds_forced(lib, rounds):
    """A small proof with the batched-affine pair rounds forced on in all five MSMs: still bit-exact vs the oracle."""
    names = ("msm_rounds", "prover_rounds_w", "prover_rounds_h")
    try:
        for nm in names:
            api.tuning_set(nm, rounds)
        c = tiny_case(31 + rounds, 700, 13, 30)
        zkd = formats.read_zkey(c["zkey_bytes"])
        exp, pub, parts = oprover.prove(zkd, c["witness"], 77, 99, return_parts=True)
        with api.Zkey(c["zkey_bytes"]) as zk, api.Prover(zk) as pr:
            got = pr.prove(c["wtns_bytes"], r=77, s=99, debug=True)
        assert got["msm_a"] == g1_plain_bytes(parts["A"])
        assert got["msm_b2"] == g2_plain_bytes(parts["B2"])
        assert got["msm_c"] == g1_plain_bytes(parts["C"])
        assert got["msm_h"] == g1_plain_bytes(parts["H"])
        assert got["proof"] == oprover.proof_to_bytes(exp)
    finally:
        for nm in names:
            api.tuning_set(nm, -1)


def test_groth16_prove_thread_safety_and_bounded_cache(lib):
    """groth16.prove keeps ONE prover per resident key: concurrent callers are serialised by the entry's lock (ctypes drops
    the GIL inside nzcp_prove), a first-call race loads the key once, in-memory keys are identified by content (a mutated
    bytearray is a different key), and the cache holds at most MAX_CACHED_KEYS keys."""
    from concurrent.futures import ThreadPoolExecutor
    groth16.terminate()
    c = tiny_case(seed=41, n_constraints=1500, n_public=4, n_free=30)
    zb = bytearray(c["zkey_bytes"])
    zmem, wmem = {"type": "mem", "data": zb}, {"type": "mem", "data": c["wtns_bytes"]}
    r, s = 0x55AA55AA, 0x77CC77CC
    want = oprover.proof_to_json(oprover.prove(formats.read_zkey(c["zkey_bytes"]), c["witness"], r, s)[0])
    with ThreadPoolExecutor(max_workers=6) as ex:
        outs = list(ex.map(lambda _: groth16.prove(zmem, wmem, r=r, s=s), range(18)))
    assert all(o["proof"] == want for o in outs)
    assert len(groth16._cache) == 1
    # same object, different content -> a different resident key (here: the same circuit set up with other toxic waste,
    # written into the same bytearray; the witness is unchanged)
    c2 = tiny_case(seed=41, n_constraints=1500, n_public=4, n_free=30, toxic=dict(TOXIC, delta=0xD1FFE7E27, tau=0x7A07A0))
    assert len(c2["zkey_bytes"]) == len(zb) and c2["zkey_bytes"] != bytes(zb)
    zb[:] = c2["zkey_bytes"]
    out2 = groth16.prove(zmem, {"type": "mem", "data": c2["wtns_bytes"]}, r=r, s=s)
    want2 = oprover.proof_to_json(oprover.prove(formats.read_zkey(c2["zkey_bytes"]), c2["witness"], r, s)[0])
    assert out2["proof"] == want2 and len(groth16._cache) == 2
    for seed in range(50, 50 + groth16.MAX_CACHED_KEYS + 1):
        ck = tiny_case(seed=seed, n_constraints=60, n_public=2, n_free=5)
        groth16.prove({"type": "mem", "data": ck["zkey_bytes"]}, {"type": "mem", "data": ck["wtns_bytes"]})
    assert len(groth16._cache) == groth16.MAX_CACHED_KEYS
    groth16.terminate()
    assert len(groth16._cache) == 0
