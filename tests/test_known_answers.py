"""Known-answer vectors from OUTSIDE this repository for the curve layer under the prover.

The reference holds no golden vector for its proving path ("parity unpinned"), but BN254 itself has
public ones: the alt_bn128 points used by the Ethereum precompile tests (EIP-196 ecAdd / ecMul: 2*G1, 3*G1;
EIP-197 / py_ecc bn128: the G2 generator and 2*G2).  snarkjs's `bn128` curve is this curve with these generators, so
every implementation here -- the Python oracle, the C oracle, the CPU verifier, and the CUDA MSM through the C ABI --
must reproduce them.  This pins field, group law and the generator encoding independently of our own restatements."""
import pytest

from nzcp_circom_b200 import verifier
from oracle import bn254 as ob
from oracle import cref
from util import g1_plain_bytes, g2_plain_bytes, le32

G1_2 = (1368015179489954701390400359078579693043519447331113978918064868415326638035,
        9918110051302171585080402603319702774565515993150576347155970296011118125764)
G1_3 = (3353031288059533942658390886683067124040920775575537747144343083137631628272,
        19321533766552368860946552437480515441416830039777911637913418824951667761761)
G2_1 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
         11559732032986387107991004021392285783925812861821192530917403151452391805634),
        (8495653923123431417604973247489272438418190587263600148770280649306958101930,
         4082367875863433681332203403145435568316851327593401208105741076214120093531))
G2_2 = ((18029695676650738226693292988307914797657423701064905010927197838374790804409,
         14583779054894525174450323658765874724019480979794335525732096752006891875705),
        (2140229616977736810657479771656733941598412651537078903776637920509952744750,
         11474861747383700316476719153975578001603231366361248090558603872215261634898))
G = (1, 2)


def test_python_oracle():
    assert ob.G2_GEN == G2_1 and ob.G1_GEN == G
    assert ob.G1.mul(G, 2) == G1_2 and ob.G1.mul(G, 3) == G1_3 and ob.G1.add(G1_2, G) == G1_3
    assert ob.G2.mul(G2_1, 2) == G2_2 and ob.G2.add(G2_1, G2_1) == G2_2
    assert ob.G1.mul(G, ob.R_MOD) is None and ob.G2.mul(G2_1, ob.R_MOD) is None
    assert ob.G1.mul(G, ob.R_MOD - 1) == (1, ob.Q_MOD - 2)


def test_c_oracle():
    g1, g2 = ob.g1_to_bytes_mont(G), ob.g2_to_bytes_mont(G2_1)
    assert cref.msm(g1, le32(2), 1, False, 1) == g1_plain_bytes(G1_2)
    assert cref.msm(g1 + g1, le32(1) + le32(2), 2, False, 1) == g1_plain_bytes(G1_3)
    assert cref.msm(g2, le32(2), 1, True, 1) == g2_plain_bytes(G2_2)


def test_cpu_verifier_group_law_and_pairing():
    assert verifier.g1_mul(G, 2) == G1_2 and verifier.g1_mul(G, 3) == G1_3 and verifier.g2_mul(G2_1, 2) == G2_2
    assert verifier.g1_on_curve(G1_3) and verifier.g2_on_curve(G2_2)
    # bilinearity on published points only: e(2 G1, G2) * e(-G1, 2 G2) = 1
    assert verifier.pairing_product_is_one([(G1_2, G2_1), (verifier.g1_neg(G), G2_2)])
    assert not verifier.pairing_product_is_one([(G1_3, G2_1), (verifier.g1_neg(G), G2_2)])


@pytest.mark.gpu
def test_gpu_msm(lib):
    from nzcp_circom_b200 import api
    g1, g2 = ob.g1_to_bytes_mont(G), ob.g2_to_bytes_mont(G2_1)
    assert api.msm(g1, le32(2), 1, g2=False)[0] == g1_plain_bytes(G1_2)
    assert api.msm(g1 + g1, le32(1) + le32(2), 2, g2=False)[0] == g1_plain_bytes(G1_3)
    assert api.msm(g2, le32(2), 1, g2=True)[0] == g2_plain_bytes(G2_2)
    # r - 1 times the generator is its negative: (1, q - 2)
    assert api.msm(g1, le32(ob.R_MOD - 1), 1, g2=False)[0] == g1_plain_bytes((1, ob.Q_MOD - 2))
