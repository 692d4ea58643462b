"""CPU: the pairing used by groth16.verify -- bilinearity, non-degeneracy, and rejection of tampered proofs."""
from nzcp_circom_b200 import verifier as v
from oracle import bn254 as ob

G1, G2 = ob.G1_GEN, ob.G2_GEN


def test_bilinear():
    e = v.pairing(G1, G2)
    assert e != v.F12_ONE
    a, b = 0x1234567, 0xFEDCBA98
    assert v.pairing(v.g1_mul(G1, a), v.g2_mul(G2, b)) == v.f12_pow(e, a * b)
    assert v.pairing(v.g1_mul(G1, a), G2) == v.pairing(G1, v.g2_mul(G2, a))
    assert v.f12_pow(e, v.R) == v.F12_ONE


def test_product_check():
    a = 987654321
    assert v.pairing_product_is_one([(v.g1_neg(v.g1_mul(G1, a)), G2), (G1, v.g2_mul(G2, a))])
    assert not v.pairing_product_is_one([(v.g1_neg(v.g1_mul(G1, a)), G2), (G1, v.g2_mul(G2, a + 1))])


def test_curve_membership():
    assert v.g1_on_curve(G1) and v.g2_on_curve(G2)
    assert not v.g1_on_curve((1, 3))
    assert v.g2_mul(G2, v.R) is None and v.g1_mul(G1, v.R) is None


def _fq2_sqrt(a):
    """Square root in Fq2 = Fq[u]/(u^2+1), q = 3 mod 4 (complex method); None if `a` is not a square."""
    q = v.Q
    a0, a1 = a
    if a1 == 0:
        s = pow(a0, (q + 1) // 4, q)
        if s * s % q == a0:
            return (s, 0)
        s = pow(-a0 % q, (q + 1) // 4, q)          # sqrt(-a0) * u
        return (0, s) if s * s % q == -a0 % q else None
    norm = (a0 * a0 + a1 * a1) % q
    s = pow(norm, (q + 1) // 4, q)
    if s * s % q != norm:
        return None
    for sg in (s, -s % q):
        t = (a0 + sg) * pow(2, -1, q) % q
        x0 = pow(t, (q + 1) // 4, q)
        if x0 * x0 % q == t and x0:
            x1 = a1 * pow(2 * x0, -1, q) % q
            if v.f2_sqr((x0, x1)) == (a0, a1):
                return (x0, x1)
    return None


def _twist_point_outside_g2():
    """A point of E'(Fq2) that is NOT in the order-r subgroup (the twist's cofactor is 2q - r)."""
    x = (5, 1)
    while True:
        y = _fq2_sqrt(v.f2_add(v.f2_mul(v.f2_sqr(x), x), v._B2))
        if y is not None:
            P = (x, y)
            assert v.g2_on_curve(P)
            if v.g2_mul(P, v.R) is not None:
                return P
        x = (x[0] + 1, x[1])


def _toy_instance():
    """A hand-made Groth16 instance with one public input, small enough for the Python verifier: vk from toxic waste,
    proof for the statement pub * 1 = pub built by the toxic-waste closed form of A, B, C."""
    from oracle import setup
    from oracle.bn254 import R_MOD
    import random
    rng = random.Random(4)
    cons, n_vars, defines = setup.random_circuit(rng, 6, 1, 2)
    tox = {"tau": 11111, "alpha": 222, "beta": 3333, "gamma": 44, "delta": 555}
    wit = setup.solve_witness(cons, n_vars, defines, 1, [3, 5])
    proof = setup.expected_proof(cons, n_vars, 1, tox, wit, 7, 9)
    zk = setup.make_zkey(cons, n_vars, 1, tox)
    from oracle import formats, prover as op
    from nzcp_circom_b200 import groth16
    vk = groth16.exportVerificationKey(formats.write_zkey(zk))
    return vk, [str(wit[1])], op.proof_to_json(proof), R_MOD


def test_verify_rejects_malformed_points():
    vk, pubs, proof, _ = _toy_instance()
    assert v.verify(vk, pubs, proof)
    # non-canonical coordinate: x + q encodes the same point mod q and must NOT be accepted
    bad = dict(proof, pi_a=[str(int(proof["pi_a"][0]) + v.Q), proof["pi_a"][1], "1"])
    assert not v.verify(vk, pubs, bad)
    bad = dict(proof, pi_c=[proof["pi_c"][0], str(int(proof["pi_c"][1]) + v.Q), "1"])
    assert not v.verify(vk, pubs, bad)
    # projective (z != 1) encodings are not part of the snarkjs proof format
    bad = dict(proof, pi_a=[proof["pi_a"][0], proof["pi_a"][1], "2"])
    assert not v.verify(vk, pubs, bad)
    # pi_b on the twist but outside the order-r subgroup
    P = _twist_point_outside_g2()
    bad = dict(proof, pi_b=[[str(P[0][0]), str(P[0][1])], [str(P[1][0]), str(P[1][1])], ["1", "0"]])
    assert not v.verify(vk, pubs, bad)
    # wrong public input / tampered point still rejected, infinity encodings parse
    assert not v.verify(vk, [str(int(pubs[0]) + 1)], proof)
    assert not v.verify(vk, pubs, dict(proof, pi_c=["0", "1", "0"]))
    assert not v.verify(vk, pubs, dict(proof, pi_b=[["0", "0"], ["1", "0"], ["0", "0"]]))
    # a verification key whose delta is off the subgroup is refused outright
    vk_bad = dict(vk, vk_delta_2=[[str(P[0][0]), str(P[0][1])], [str(P[1][0]), str(P[1][1])], ["1", "0"]])
    assert not v.verify(vk_bad, pubs, proof)
    assert v.verify(vk, pubs, proof)
