"""groth16.fullProve(input, wasmFile, zkey) with a .wasm PATH, as the reference would call it:
the witness step is delegated to the reference's own toolchain (node + snarkjs `wtns.calculate`) in a child process.
There is no node in this environment, so the child is a stand-in executable that honours the same command line
(`node -e <script> input.json circuit.wasm out.wtns`): what is tested is everything on OUR side of that boundary -- the
input JSON handed over, the .wtns picked up, error propagation, and the clear failure when no node exists."""
import json
import os
import stat
import sys

import pytest

from nzcp_circom_b200 import groth16
from nzcp_circom_b200._lib import NzcpError
from util import tiny_case


def _fake_node(tmp_path, wtns_bytes, fail=False):
    """An executable that behaves like `node -e SCRIPT in.json x.wasm out.wtns` for snarkjs wtns.calculate."""
    golden = tmp_path / "golden.wtns"
    golden.write_bytes(wtns_bytes)
    seen = tmp_path / "seen.json"
    exe = tmp_path / "node"
    exe.write_text(
        "#!%s\n"
        "import json, shutil, sys\n"
        "assert sys.argv[1] == '-e' and 'wtns.calculate' in sys.argv[2]\n"
        "inp, wasm, out = sys.argv[3:6]\n"
        "json.dump({'input': json.load(open(inp)), 'wasm': wasm}, open(%r, 'w'))\n"
        "if %r:\n"
        "    sys.stderr.write('Error: Assert Failed. Error in template NZCPPubIdentity_1 line: 470\\n'); sys.exit(1)\n"
        "shutil.copy(%r, out)\n" % (sys.executable, str(seen), fail, str(golden)))
    exe.chmod(exe.stat().st_mode | stat.S_IEXEC)
    return str(exe), seen


def test_calculate_witness_through_node_child(tmp_path):
    c = tiny_case(seed=3, n_constraints=30, n_public=2, n_free=5)
    node, seen = _fake_node(tmp_path, c["wtns_bytes"])
    wasm = tmp_path / "nzcp_exampleTest.wasm"
    wasm.write_bytes(b"\0asm\x01\0\0\0")
    inp = {"toBeSigned": [1, 0, 1, 1], "toBeSignedLen": 314, "big": 2 ** 200}
    got = groth16.calculate_witness(inp, str(wasm), node=node)
    assert got == c["wtns_bytes"]
    rec = json.load(open(seen))
    assert rec["input"] == {"toBeSigned": [1, 0, 1, 1], "toBeSignedLen": 314, "big": str(2 ** 200)}   # > 2^53: as a string
    assert rec["wasm"] == os.path.abspath(str(wasm))
    # the circuit's own assertion failures come back as the error text
    node_bad, _ = _fake_node(tmp_path, c["wtns_bytes"], fail=True)
    with pytest.raises(NzcpError) as e:
        groth16.calculate_witness(inp, str(wasm), node=node_bad)
    assert "Assert Failed" in str(e.value)
    with pytest.raises(NzcpError) as e:
        groth16.calculate_witness(inp, str(tmp_path / "missing.wasm"), node=node)
    assert "wasm file not found" in str(e.value)


def test_fullprove_without_node_fails_loudly(tmp_path, monkeypatch):
    wasm = tmp_path / "c.wasm"
    wasm.write_bytes(b"\0asm\x01\0\0\0")
    monkeypatch.delenv("NZCP_NODE", raising=False)
    monkeypatch.setenv("PATH", str(tmp_path))          # no node anywhere
    with pytest.raises(NzcpError) as e:
        groth16.fullProve({"a": 1}, str(wasm), b"zkey")
    assert "node" in str(e.value) and e.value.code == -1


@pytest.mark.gpu
def test_fullprove_with_wasm_path_end_to_end(lib, tmp_path):
    from oracle import formats
    from oracle import prover as oprover
    c = tiny_case(seed=8, n_constraints=300, n_public=4, n_free=9)
    node, _ = _fake_node(tmp_path, c["wtns_bytes"])
    wasm = tmp_path / "circuit.wasm"
    wasm.write_bytes(b"\0asm\x01\0\0\0")
    zkey = tmp_path / "circuit_final.zkey"
    zkey.write_bytes(c["zkey_bytes"])
    out = groth16.fullProve({"in": [1, 2, 3]}, str(wasm), str(zkey), r=11, s=13, node=node)
    exp, pub = oprover.prove(formats.read_zkey(c["zkey_bytes"]), c["witness"], 11, 13)
    assert out["proof"] == oprover.proof_to_json(exp) and out["publicSignals"] == [str(x) for x in pub]
    assert groth16.verify(groth16.exportVerificationKey(str(zkey)), out["publicSignals"], out["proof"])
    groth16.terminate()
