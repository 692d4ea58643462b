"""CPU: pin the oracle.  The reference holds no golden vector for the prover (snarkjs is an
un-vendored dependency with no call site), so the oracle is pinned by mathematics instead:
  * constants against SURVEY.md 8c (computed independently with sympy there),
  * NTT against the O(n^2) definition,
  * multiexp against naive sum of scalar multiples,
  * the whole prover against the toxic-waste closed form and the pairing equation,
  * the committed golden fixtures (tests/golden/) regenerate byte-identically.
"""
import random

import pytest

from nzcp_circom_b200 import groth16, verifier
from oracle import bn254 as ob
from oracle import formats, setup
from oracle import prover as oprover
from util import tiny_case

R, Q = ob.R_MOD, ob.Q_MOD


def test_constants_survey_8c():
    assert R == 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    assert Q == 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    assert (R - 1) % (1 << 28) == 0 and (R - 1) % (1 << 29) != 0
    assert ob.FR_W[28] == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert ob.FR_W[20] == 17220337697351015657950521176323262483320249231368149235373741788599650842711
    assert ob.FR_W[21] == 13536764371732269273912573961853310557438878140379554347802702086337840854307
    assert (1 << 256) % R == 6350874878119819312338956282401532410528162663560392320966563075034087161851
    assert pow(1 << 256, 2, R) == 944936681149208446651664254269745548490766851729442924617792859073125903783
    assert (1 << 256) % Q == 6350874878119819312338956282401532409788428879151445726012394534686998597021
    assert pow(1 << 256, 2, Q) == 3096616502983703923843567936837374451735540968419076528771170197431451843209
    # on the odd coset Z(x) = x^n - 1 is the constant -2 (why snarkjs never divides by Z)
    assert pow(ob.FR_W[21], 1 << 20, R) == R - 1
    # group orders
    assert ob.G1.mul(ob.G1_GEN, R) is None and ob.G2.mul(ob.G2_GEN, R) is None


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 6])
def test_ntt_matches_definition(log_n):
    rng = random.Random(log_n)
    n = 1 << log_n
    a = [rng.randrange(R) for _ in range(n)]
    w = ob.FR_W[log_n]
    naive = [sum(a[j] * pow(w, j * k, R) for j in range(n)) % R for k in range(n)]
    assert ob.ntt(a) == naive
    assert ob.ntt(naive, inverse=True) == a


@pytest.mark.parametrize("g2", [False, True])
def test_multiexp_matches_naive(g2):
    rng = random.Random(9 + g2)
    curve = ob.G2 if g2 else ob.G1
    pts = [curve.mul(curve.gen, rng.randrange(1, 1000)) for _ in range(20)] + [None]
    sc = [rng.choice([0, 1, rng.randrange(R), R - 1]) for _ in range(21)]
    naive = None
    for P, k in zip(pts, sc):
        naive = curve.add(naive, curve.mul(P, k) if P is not None else None)
    assert curve.to_affine(oprover.multiexp(curve, pts, sc)) == naive


@pytest.mark.parametrize("seed,nc,npub,nfree", [(1, 1, 1, 4), (2, 6, 2, 3), (3, 28, 3, 5), (4, 29, 3, 5), (5, 100, 0, 8)])
def test_oracle_prover_equals_closed_form_and_verifies(seed, nc, npub, nfree):
    c = tiny_case(seed, nc, npub, nfree)
    rng = random.Random(seed)
    r, s = rng.randrange(R), rng.randrange(R)
    zk = formats.read_zkey(c["zkey_bytes"])
    proof, pub = oprover.prove_files(c["zkey_bytes"], c["wtns_bytes"], r, s)
    assert proof == setup.expected_proof(c["constraints"], c["n_vars"], npub, c["toxic"], c["witness"], r, s)
    assert pub == c["witness"][1:npub + 1]
    vk = groth16.exportVerificationKey(c["zkey_bytes"])
    assert verifier.verify(vk, [str(x) for x in pub], oprover.proof_to_json(proof))
    # format self-check of SURVEY.md 7: the last nPublic+1 coefficient records are exactly R^2 mod r
    raw = zk["coefs_raw_section"]
    r2 = pow(1 << 256, 2, R)
    for i in range(npub + 1):
        o = len(raw) - 44 * (npub + 1 - i)
        assert int.from_bytes(raw[o + 12:o + 44], "little") == r2
    for P in zk["A"] + zk["B1"] + zk["C"] + zk["H"] + zk["IC"]:
        assert ob.G1.on_curve(P)
    for P in zk["B2"]:
        assert ob.G2.on_curve(P)


def test_h_is_zero_polynomial_quotient_times_minus_two():
    """F6 in SURVEY.md: the h scalars are evaluations of A*B-C on the odd coset, where Z = -2; so h / -2 are the
    evaluations of the quotient polynomial H, which must have degree < n - 1."""
    c = tiny_case(8, 13, 2, 4)
    zk = formats.read_zkey(c["zkey_bytes"])
    h = oprover.h_scalars(zk, c["witness"])
    n = zk["domainSize"]
    lg = n.bit_length() - 1
    inc = ob.FR_W[lg + 1]
    minus_half = pow(R - 2, -1, R)
    coef = ob.ntt([x * minus_half % R for x in h], inverse=True)
    coef = [x * pow(inc, -i, R) % R for i, x in enumerate(coef)]       # undo the coset shift
    assert coef[-1] == 0                                                # deg H <= n - 2
    # and H(tau) * Z(tau) == A(tau) B(tau) - C(tau)
    tau = c["toxic"]["tau"]
    _, At, Bt, Ct = setup.poly_evals(c["constraints"], c["n_vars"], 2, tau)
    a, b, cc = (sum(w * x for w, x in zip(c["witness"], v)) % R for v in (At, Bt, Ct))
    Ht = sum(x * pow(tau, i, R) for i, x in enumerate(coef)) % R
    assert Ht * (pow(tau, n, R) - 1) % R == (a * b - cc) % R


def test_container_roundtrip_and_errors():
    c = tiny_case(6, 5, 1, 3)
    assert formats.read_wtns(c["wtns_bytes"])["witness"] == c["witness"]
    zk = formats.read_zkey(c["zkey_bytes"])
    assert formats.write_zkey(zk) == c["zkey_bytes"]
    with pytest.raises(ValueError, match="Invalid witness length"):
        oprover.prove(zk, c["witness"][:-1], 1, 1)
    r1 = formats.write_r1cs(c["n_vars"], 1, 0, 3, c["constraints"])
    back = formats.read_r1cs(r1)
    assert back["constraints"] == [tuple({k: v % R for k, v in lc.items()} for lc in con) for con in c["constraints"]]
