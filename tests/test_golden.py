"""Golden fixtures (tests/golden/, made by make_golden.py with the pure-Python oracle):
CPU -- both oracle restatements reproduce them and the generator is deterministic;
GPU -- the CUDA path reproduces every intermediate and the proof, through the C ABI."""
import json
import os

import pytest

from oracle import cref, formats
from oracle import prover as oprover

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["g32", "g256"]


def load(name):
    z = open(os.path.join(HERE, name + ".zkey"), "rb").read()
    w = open(os.path.join(HERE, name + ".wtns"), "rb").read()
    return z, w, json.load(open(os.path.join(HERE, name + ".json")))


@pytest.mark.parametrize("name", NAMES)
def test_python_oracle_reproduces_golden(name):
    z, w, exp = load(name)
    proof, pub, parts = oprover.prove_files(z, w, int(exp["r"]), int(exp["s"]), return_parts=True)
    assert oprover.proof_to_bytes(proof).hex() == exp["proof_hex"]
    assert oprover.proof_to_json(proof) == exp["proof"]
    assert [str(x) for x in pub] == exp["publicSignals"]
    assert b"".join(int(x).to_bytes(32, "little") for x in parts["h"]).hex() == exp["h_hex"]


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_reproduces_golden(name):
    z, w, exp = load(name)
    got = cref.prove(z, w, int(exp["r"]), int(exp["s"]), threads=3, want_h_size=exp["domain_size"])
    assert got["proof"].hex() == exp["proof_hex"]
    assert got["h"].hex() == exp["h_hex"]
    for k in ("msm_a", "msm_b1", "msm_b2", "msm_c", "msm_h"):
        assert got[k].hex() == exp[k], k


def test_generator_is_deterministic():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    z, w, exp = mod.build("g32")
    z0, w0, exp0 = load("g32")
    assert z == z0 and w == w0 and exp == exp0
    hdr, _ = formats.read_zkey_header(z0)
    assert (hdr["nVars"], hdr["nPublic"], hdr["domainSize"]) == (exp0["n_vars"], exp0["n_public"], exp0["domain_size"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_golden(lib, name):
    from nzcp_circom_b200 import api, groth16
    z, w, exp = load(name)
    with api.Zkey(z) as zk, api.Prover(zk) as pr:
        got = pr.prove(w, r=int(exp["r"]), s=int(exp["s"]), debug=True, want_h=True)
    assert got["h"].hex() == exp["h_hex"]
    for k in ("msm_a", "msm_b1", "msm_b2", "msm_c", "msm_h"):
        assert got[k].hex() == exp[k], k
    assert got["proof"].hex() == exp["proof_hex"]
    out = groth16.prove(os.path.join(HERE, name + ".zkey"), os.path.join(HERE, name + ".wtns"),
                        r=int(exp["r"]), s=int(exp["s"]))
    assert out["proof"] == exp["proof"] and out["publicSignals"] == exp["publicSignals"]
    groth16.terminate()
