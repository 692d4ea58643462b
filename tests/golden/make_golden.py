"""Generates the committed golden fixtures with the pure-Python ORACLE (oracle/*.py).

    python tests/golden/make_golden.py

The reference (noway/nzcp-circom) holds no prover vector and snarkjs cannot run here, so these are
vectors of OUR restatement of snarkjs groth16.prove, cross-checked at generation time against the toxic-waste closed
form and the pairing check.  They pin the oracle against regressions and give the GPU tests a fixed target.
Each case: <name>.zkey, <name>.wtns (snarkjs binary formats) and <name>.json (r, s, h scalars, the five MSM results,
the proof; hex little-endian / decimal strings).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from nzcp_circom_b200 import groth16, verifier  # noqa: E402
from oracle import formats, setup  # noqa: E402
from oracle import prover as oprover  # noqa: E402
from util import g1_plain_bytes, g2_plain_bytes, tiny_case  # noqa: E402

CASES = {
    # name: (seed, n_constraints, n_public, n_free, r, s)
    "g32": (101, 28, 3, 5, 0x1B2C3D4E5F60718293A4B5C6D7E8F9, 0x0F1E2D3C4B5A69788796A5B4C3D2E1F0),
    "g256": (102, 200, 5, 12, 0x123456789ABCDEF0FEDCBA9876543210 << 90, (1 << 253) - 12345),
}


def build(name):
    seed, nc, npub, nfree, r, s = CASES[name]
    c = tiny_case(seed, nc, npub, nfree)
    zk = formats.read_zkey(c["zkey_bytes"])
    proof, pub, parts = oprover.prove(zk, c["witness"], r, s, return_parts=True)
    assert proof == setup.expected_proof(c["constraints"], c["n_vars"], npub, c["toxic"], c["witness"], r, s)
    pj = oprover.proof_to_json(proof)
    assert verifier.verify(groth16.exportVerificationKey(c["zkey_bytes"]), [str(x) for x in pub], pj)
    exp = {
        "n_vars": c["n_vars"], "n_public": npub, "domain_size": zk["domainSize"], "r": str(r), "s": str(s),
        "h_hex": b"".join(int(x).to_bytes(32, "little") for x in parts["h"]).hex(),
        "msm_a": g1_plain_bytes(parts["A"]).hex(), "msm_b1": g1_plain_bytes(parts["B1"]).hex(),
        "msm_b2": g2_plain_bytes(parts["B2"]).hex(), "msm_c": g1_plain_bytes(parts["C"]).hex(),
        "msm_h": g1_plain_bytes(parts["H"]).hex(),
        "proof_hex": oprover.proof_to_bytes(proof).hex(), "proof": pj, "publicSignals": [str(x) for x in pub],
    }
    return c["zkey_bytes"], c["wtns_bytes"], exp


if __name__ == "__main__":
    for name in CASES:
        z, w, exp = build(name)
        open(os.path.join(HERE, name + ".zkey"), "wb").write(z)
        open(os.path.join(HERE, name + ".wtns"), "wb").write(w)
        json.dump(exp, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
        print(name, len(z), len(w))
