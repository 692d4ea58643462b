"""CPU: prover input and expected public signals for the NZ COVID Pass spec example (the one reproducible pass of the
reference's suite: EXAMPLE_PASS_URI at /root/reference/test/nzcp.js:51; its live passes need secrets).

Golden values: SURVEY.md section 4 (derived there by following the reference's helpers) and the offsets the reference's
own tests assert -- FindVCAndExp -> (76, 68) at test/nzcp.js:89, FindCredSubj -> 246 at :144."""
import os

import pytest

from nzcp_circom_b200 import nzcp_input as ni

URI = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "example_pass_uri.txt")).read().strip()


def test_to_be_signed_matches_reference_offsets():
    tbs = ni.to_be_signed(URI)
    assert len(tbs) == 314                                  # exactly MaxToBeSignedBytes of nzcp_exampleTest (test/nzcp.js:314)
    assert tbs[:12] == b"\x84\x6aSignature1"
    assert tbs[27] == 0xA5                                  # the CWT claims map
    assert tbs[68:73] == bytes.fromhex("1a7450400a")        # exp claim at 68: uint32 1951416330
    assert tbs[76] == 0xA4 and tbs[246] == 0xA3            # vc map at 76, credentialSubject map at 246
    r, s = ni.signature_rs(URI)
    assert len(r) == len(s) == 64


def test_public_identity_golden():
    pid = ni.public_identity(URI)
    assert pid["credSubjConcat"] == "Jack,Sparrow,1960-04-16"
    assert pid["credSubjHash"] == "5fb355822221720ea4ce6734e5a09e459d452574a19310c0cea7c141f43a3dab"
    assert pid["toBeSignedHash"] == "271ce33d671a2d3b816d788135f4343e14bc66802f8cd841faac939e8c11f3ee"
    assert pid["exp"] == 1951416330


def test_circuit_input_and_public_signals_layout():
    inp = ni.circuit_input(URI, 314)
    assert inp["toBeSignedLen"] == 314 and len(inp["toBeSigned"]) == 314 * 8
    assert inp["toBeSigned"][:8] == [1, 0, 0, 0, 0, 1, 0, 0]          # 0x84, MSB first (helpers/utils.js:2-10)
    live = ni.circuit_input(URI, 355)                                  # nzcp_liveTest shape: zero padded to 355 bytes
    assert live["toBeSignedLen"] == 314 and len(live["toBeSigned"]) == 355 * 8 and not any(live["toBeSigned"][314 * 8:])
    with pytest.raises(ni.PassError):
        ni.circuit_input(URI, 313)
    sig = ni.expected_public_signals(URI)
    assert len(sig) == 513 and sig[512] == "1951416330"
    # witness.slice(1, 257) packed MSB-first is the credSubj hash (test/nzcp.js:41-42), the next 256 the ToBeSigned hash
    def pack(bits):
        return bytes(sum(int(b) << (7 - j) for j, b in enumerate(bits[i:i + 8])) for i in range(0, len(bits), 8)).hex()
    assert pack(sig[:256]) == "5fb355822221720ea4ce6734e5a09e459d452574a19310c0cea7c141f43a3dab"
    assert pack(sig[256:512]) == "271ce33d671a2d3b816d788135f4343e14bc66802f8cd841faac939e8c11f3ee"
    assert ni.check_public_signals(URI, sig) and not ni.check_public_signals(URI, sig[:-1] + ["0"])


def test_rejects_malformed():
    with pytest.raises(ni.PassError):
        ni.decode_pass("NZCP:/2/ABC")
    with pytest.raises(ni.PassError):
        ni.decode_pass("NZCP:/1/AAAA")
    with pytest.raises(ni.PassError):
        ni.decode_pass(URI[:200])
