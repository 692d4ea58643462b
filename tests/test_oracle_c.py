"""CPU: the C restatement (oracle/c) must agree bit-for-bit with the pure-Python restatement (oracle/*.py), with the
toxic-waste closed form, and be independent of its thread count."""
import random

import pytest

from oracle import bn254 as ob
from oracle import cref, formats, setup
from oracle import prover as oprover
from util import g1_plain_bytes, g2_plain_bytes, le32, tiny_case

R, Q = ob.R_MOD, ob.Q_MOD
MONT = 1 << 256


@pytest.mark.parametrize("field,p", [(0, R), (1, Q)])
def test_c_field_ops(field, p):
    rng = random.Random(field)
    a = [0, 1, p - 1, MONT % p, p - 1] + [rng.randrange(p) for _ in range(2000)]
    b = [0, p - 1, p - 1, 1, 2] + [rng.randrange(p) for _ in range(2000)]
    ab, bb = b"".join(map(le32, a)), b"".join(map(le32, b))
    rinv = pow(MONT, -1, p)
    assert cref.field_op(field, 0, ab, bb, len(a)) == b"".join(le32(x * y * rinv % p) for x, y in zip(a, b))
    assert cref.field_op(field, 1, ab, bb, len(a)) == b"".join(le32((x + y) % p) for x, y in zip(a, b))
    assert cref.field_op(field, 2, ab, bb, len(a)) == b"".join(le32((x - y) % p) for x, y in zip(a, b))


@pytest.mark.parametrize("log_n", [0, 1, 4, 9])
def test_c_ntt_vs_python(log_n):
    rng = random.Random(log_n)
    n = 1 << log_n
    a = [rng.randrange(R) for _ in range(n)]
    rinv = pow(MONT, -1, R)

    def enc(v):
        return bytearray(b"".join(le32(x * MONT % R) for x in v))

    def dec(buf):
        return [int.from_bytes(buf[i:i + 32], "little") * rinv % R for i in range(0, len(buf), 32)]
    for inv in (False, True):
        for th in (1, 3):
            buf = enc(a)
            cref.ntt(buf, log_n, inv, th)
            assert dec(buf) == ob.ntt(a, inverse=inv)
    buf = enc(a)
    cref.ntt_coset(buf, log_n, 2)
    inc = ob.FR_W[log_n + 1]
    sc = [x * pow(inc, i, R) % R for i, x in enumerate(ob.ntt(a, inverse=True))]
    assert dec(buf) == ob.ntt(sc)


@pytest.mark.parametrize("g2", [False, True])
def test_c_msm_vs_python(g2):
    rng = random.Random(3 + g2)
    curve = ob.G2 if g2 else ob.G1
    fb = ob.FixedBase(curve, curve.gen)
    P = fb.mul(5)
    pts = [P, P, curve.neg(P), None] + [fb.mul(rng.randrange(1, R)) for _ in range(60)]
    sc = [3, 3, 3, 9] + [rng.choice([0, 1, R - 1, rng.randrange(256), rng.randrange(R)]) for _ in range(60)]
    enc = ob.g2_to_bytes_mont if g2 else ob.g1_to_bytes_mont
    exp = curve.to_affine(oprover.multiexp(curve, pts, sc))
    want = g2_plain_bytes(exp) if g2 else g1_plain_bytes(exp)
    for th in (1, 4):
        assert cref.msm(b"".join(enc(p) for p in pts), b"".join(map(le32, sc)), len(pts), g2, th) == want
    assert cref.msm(b"", b"", 0, g2, 1) == bytes(128 if g2 else 64)


@pytest.mark.parametrize("seed,nc,npub,nfree", [(1, 1, 1, 4), (3, 28, 3, 5), (4, 29, 3, 5), (5, 300, 6, 9)])
def test_c_prover_vs_python_and_closed_form(seed, nc, npub, nfree):
    c = tiny_case(seed, nc, npub, nfree)
    r, s = random.Random(seed).randrange(R), random.Random(seed + 1).randrange(R)
    zk = formats.read_zkey(c["zkey_bytes"])
    exp, _pub, parts = oprover.prove(zk, c["witness"], r, s, return_parts=True)
    for th in (1, 5):
        got = cref.prove(c["zkey_bytes"], c["wtns_bytes"], r, s, threads=th, want_h_size=zk["domainSize"])
        assert got["h"] == b"".join(map(le32, parts["h"]))
        assert got["msm_a"] == g1_plain_bytes(parts["A"]) and got["msm_b1"] == g1_plain_bytes(parts["B1"])
        assert got["msm_b2"] == g2_plain_bytes(parts["B2"])
        assert got["msm_c"] == g1_plain_bytes(parts["C"]) and got["msm_h"] == g1_plain_bytes(parts["H"])
        assert got["proof"] == oprover.proof_to_bytes(exp)
    cf = setup.expected_proof(c["constraints"], c["n_vars"], npub, c["toxic"], c["witness"], r, s)
    assert got["proof"] == oprover.proof_to_bytes(cf)


def test_c_prover_errors():
    c = tiny_case(2, 5, 1, 3)
    with pytest.raises(ValueError, match="Invalid witness length"):
        cref.prove(c["zkey_bytes"], formats.write_wtns(c["witness"][:-1]), 1, 1)
    with pytest.raises(ValueError, match="Invalid File format"):
        cref.prove(b"abcd" + c["zkey_bytes"][4:], c["wtns_bytes"], 1, 1)
