"""`zkey new` on the GPU: .r1cs + prepared .ptau -> .zkey, against the toxic-waste closed form.

The ptau is written by the oracle from a KNOWN (tau, alpha, beta) (oracle/ptau.py), so the key the library derives from
it must equal, section by section, the key oracle/setup.py computes as scalars times generators with gamma = delta = 1 --
an identity that does not depend on either side's point arithmetic agreeing by accident.  Then the derived key is used:
it passes the format self-check, a proof made with it matches the oracle's proof and verifies.
CPU part: size computation and snarkjs's three error messages (raised before any device work).
"""
import random
import struct

import pytest

from nzcp_circom_b200 import api, groth16
from nzcp_circom_b200._lib import NzcpError
from oracle import formats, ptau, setup
from oracle import prover as oprover
from oracle.bn254 import R_MOD

TOX = {"tau": 0x5EED5EED5EED, "alpha": 0xA1A1A1, "beta": 0xB2B2B2B2, "gamma": 1, "delta": 1}


def _circuit(seed, n_constraints, n_public, n_free):
    rng = random.Random(seed)
    cons, n_vars, defines = setup.random_circuit(rng, n_constraints, n_public, n_free)
    # some negative coefficients (r - 1, r - 2), as circom emits: exercises the (r - k) * (-P) folding of the kernel
    for A, B, _C in cons[:5]:
        A[next(iter(A))] = R_MOD - 1
        B[next(iter(B))] = R_MOD - 2
    free = [rng.randrange(R_MOD) if i % 3 == 0 else rng.randrange(2) for i in range(n_free)]
    wit = setup.solve_witness(cons, n_vars, defines, n_public, free)
    r1cs = formats.write_r1cs(n_vars, n_public, 0, n_vars - 1 - n_public, cons)
    return cons, n_vars, wit, r1cs


def _sections(buf):
    _, secs = formats.read_container(buf, b"zkey")
    return {sid: bytes(memoryview(buf)[v[0][0]:v[0][0] + v[0][1]]) for sid, v in secs.items()}, \
        [sid for sid, _ in sorted(((sid, v[0][0]) for sid, v in secs.items()), key=lambda x: x[1])]


@pytest.fixture(scope="module")
def ptau5():
    return ptau.write_ptau(TOX["tau"], TOX["alpha"], TOX["beta"], 5)


def test_zkey_new_size_and_errors_cpu(lib, ptau5):
    import ctypes as C
    from nzcp_circom_b200 import _lib
    cons, n_vars, _, r1cs = _circuit(1, 20, 3, 4)
    size = C.c_size_t()
    _lib.check(lib.nzcp_zkey_new_size(_lib.addr(r1cs), len(r1cs), _lib.addr(ptau5), len(ptau5), C.byref(size)))
    want = formats.write_zkey(setup.make_zkey(cons, n_vars, 3, TOX))
    ws, _ = _sections(want)
    assert size.value == 12 + 10 * 12 + sum(len(ws[k]) for k in range(1, 10)) + 68    # same sections 1-9, plus section 10
    # "Powers of tau is not prepared."
    raw = ptau.write_ptau(TOX["tau"], TOX["alpha"], TOX["beta"], 5, prepared=False)
    with pytest.raises(NzcpError) as e:
        api.zkey_new(r1cs, raw)
    assert "Powers of tau is not prepared." in str(e.value)
    # circuit too big for the ceremony: 20 + 3 -> power 5 fits, 40 + 3 -> power 6 does not
    _, _, _, big = _circuit(2, 40, 3, 4)
    with pytest.raises(NzcpError) as e:
        api.zkey_new(big, ptau5)
    assert "circuit too big for this power of tau ceremony. 40*2 > 2**5" in str(e.value)
    # r1cs over another prime
    bad = bytearray(r1cs)
    off = formats.read_container(bad, b"r1cs")[1][1][0][0]
    bad[off + 4] ^= 2
    with pytest.raises(NzcpError) as e:
        api.zkey_new(bad, ptau5)
    assert "r1cs curve does not match powers of tau ceremony curve" in str(e.value)
    with pytest.raises(NzcpError):
        api.zkey_new(b"r1cs" + bytes(40), ptau5)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,nc,npub,nfree,power", [(3, 20, 3, 4, 5), (4, 28, 3, 6, 5), (5, 5, 2, 3, 5), (6, 9, 0, 4, 5)])
def test_zkey_new_matches_closed_form(lib, ptau5, seed, nc, npub, nfree, power):
    """power 5 ceremony, circuits of power 5, 5 (exactly full: 28 + 3 + 1 = 32), 3 and 4: the level offsets inside the
    Lagrange sections are exercised below and at the ceremony's own size."""
    cons, n_vars, wit, r1cs = _circuit(seed, nc, npub, nfree)
    got = api.zkey_new(r1cs, ptau5)
    want = formats.write_zkey(setup.make_zkey(cons, n_vars, npub, TOX))
    gs, order = _sections(got)
    ws, _ = _sections(want)
    assert order == [1, 2, 4, 3, 9, 8, 5, 6, 7, 10]            # zkey_new.js writes them in this order
    for sid, name in ((1, "protocol"), (2, "header"), (4, "coefs"), (3, "IC"), (5, "A"), (6, "B1"), (7, "B2"), (8, "C"), (9, "H")):
        assert gs[sid] == ws[sid], name
    assert gs[10] == bytes(64) + struct.pack("<I", 0)
    # the derived key is a working proving key
    rep = api.zkey_selfcheck(got)
    assert rep["ok"] and rep["n_constraints"] == nc, rep
    wt = formats.write_wtns(wit)
    r, s = 0x1357, 0x2468
    out = groth16.prove({"type": "mem", "data": got}, {"type": "mem", "data": wt}, r=r, s=s)
    exp, _ = oprover.prove_files(want, wt, r, s)
    assert out["proof"] == oprover.proof_to_json(exp)
    assert groth16.verify(groth16.zKey.exportVerificationKey(got), out["publicSignals"], out["proof"])
    groth16.terminate()


@pytest.mark.gpu
def test_zkey_new_mid_size_proves_and_verifies(lib):
    """2^11 domain from the GPU-made synthetic circuit's own .r1cs: the key derived from a ptau equals the key the
    synthetic setup kernel makes from the same toxic waste (two independent GPU paths), byte for byte per section."""
    sc = api.SynthCircuit(seed=77, n_constraints=1500, n_public=9, n_free=40)
    r1cs = sc.r1cs()
    pt = ptau.write_ptau(TOX["tau"], TOX["alpha"], TOX["beta"], 11)
    got = api.zkey_new(r1cs, pt)
    want = sc.zkey([TOX[k] for k in ("tau", "alpha", "beta", "gamma", "delta")])
    gs, _ = _sections(got)
    ws, _ = _sections(want)
    for sid in range(1, 10):
        assert gs[sid] == ws[sid], sid
    wt = sc.wtns(5)
    out = groth16.prove({"type": "mem", "data": got}, {"type": "mem", "data": wt})
    assert groth16.verify(groth16.exportVerificationKey(got), out["publicSignals"], out["proof"])
    groth16.terminate()
