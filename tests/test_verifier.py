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
