"""Malformed-container fuzzing of every parser behind the C ABI (zkey, wtns, r1cs, ptau): random truncations, bit flips in
the headers / section tables / length fields and oversized counts must come back as an NZCP_E_* code (or succeed) --
never crash, hang or read out of bounds.  The parsers run before any device work, so the CPU box covers them (a
structurally valid zkey then fails with NZCP_E_CUDA for want of a device, which is fine here); the wtns reader sits
behind a prover handle and is fuzzed on the GPU box.

snarkjs throws on the same inputs (binfileutils readBinFile: "Invalid File format", "Version not supported", section
size checks); only the error CODE is asserted here, the three snarkjs message texts are pinned in test_gpu_prove.py.
"""
import ctypes as C
import random
import struct

import pytest

from nzcp_circom_b200 import _lib, api
from nzcp_circom_b200._lib import NzcpError
from oracle import formats, ptau, setup
from util import tiny_case

ALLOWED = {-1, -2, -3, -4, -5, -6, -7}      # every documented error except INTERNAL (-8): a parser must classify its input


def _mutations(buf, rng, count):
    n = len(buf)
    yield bytes(buf[:rng.randrange(0, min(n, 64))])                  # cut inside the header
    yield bytes(buf[:n - 1])
    yield bytes(buf) + b"\0" * 7
    for _ in range(count):
        b = bytearray(buf)
        kind = rng.randrange(6)
        if kind == 0:                                               # truncate anywhere
            yield bytes(b[:rng.randrange(0, n)])
            continue
        if kind == 1:                                               # flip a bit in the first 400 bytes (header, section table)
            i = rng.randrange(0, min(n, 400))
            b[i] ^= 1 << rng.randrange(8)
        elif kind == 2:                                             # poke a 32-bit field with an extreme value
            i = rng.randrange(0, min(n - 4, 600))
            struct.pack_into("<I", b, i, rng.choice([0, 1, 0x7FFFFFFF, 0xFFFFFFFF, 0x80000000, n, n + 1]))
        elif kind == 3:                                             # poke a 64-bit section length
            i = rng.randrange(0, min(n - 8, 600))
            struct.pack_into("<Q", b, i, rng.choice([0, n, n * 2, 2 ** 63, 2 ** 64 - 1, 2 ** 32 + 5]))
        elif kind == 4:                                             # random byte anywhere
            b[rng.randrange(n)] = rng.randrange(256)
        else:                                                       # swap two 12-byte windows
            i, j = rng.randrange(0, n - 12), rng.randrange(0, n - 12)
            b[i:i + 12], b[j:j + 12] = b[j:j + 12], b[i:i + 12]
        yield bytes(b)


def _call(fn):
    try:
        fn()
        return 0
    except NzcpError as e:
        return e.code


@pytest.fixture(scope="module")
def samples():
    c = tiny_case(seed=77, n_constraints=40, n_public=3, n_free=6)
    r1cs = formats.write_r1cs(c["n_vars"], 3, 0, c["n_vars"] - 4, c["constraints"])
    pt = ptau.write_ptau(11, 22, 33, 6)
    return {"zkey": bytes(c["zkey_bytes"]), "wtns": bytes(c["wtns_bytes"]), "r1cs": r1cs, "ptau": pt, "case": c}


def test_zkey_parsers_never_crash(lib, samples):
    rng = random.Random(1)
    zk = samples["zkey"]
    codes = set()
    for m in _mutations(zk, rng, 400):
        def load():
            h = C.c_void_p()
            rc = lib.nzcp_zkey_load(_lib.addr(m), len(m), 0, C.byref(h))
            if rc == 0:
                lib.nzcp_zkey_free(h)
            _lib.check(rc)
        codes.add(_call(load))
        codes.add(_call(lambda: api.zkey_selfcheck(m)))
    assert codes <= ALLOWED | {0}, codes
    assert -2 in codes                                   # malformed containers were seen and classified as such


def test_r1cs_and_ptau_parsers_never_crash(lib, samples):
    rng = random.Random(2)
    r1cs, pt = samples["r1cs"], samples["ptau"]
    size = C.c_size_t()
    codes = set()
    for m in _mutations(r1cs, rng, 300):
        codes.add(_call(lambda: _lib.check(lib.nzcp_zkey_new_size(_lib.addr(m), len(m), _lib.addr(pt), len(pt), C.byref(size)))))
    for m in _mutations(pt, rng, 300):
        codes.add(_call(lambda: _lib.check(lib.nzcp_zkey_new_size(_lib.addr(r1cs), len(r1cs), _lib.addr(m), len(m), C.byref(size)))))
    assert codes <= ALLOWED | {0}, codes
    assert -2 in codes


@pytest.mark.gpu
def test_wtns_reader_never_crashes_and_valid_inputs_still_prove(lib, samples):
    rng = random.Random(3)
    c = samples["case"]
    with api.Zkey(samples["zkey"]) as zk, api.Prover(zk) as pr:
        good = pr.prove(samples["wtns"], r=5, s=6)["proof"]
        codes = set()
        for m in _mutations(samples["wtns"], rng, 300):
            codes.add(_call(lambda: pr.prove(m, r=5, s=6)))
        assert codes <= ALLOWED | {0}, codes
        assert {-2, -5} & codes
        assert pr.prove(samples["wtns"], r=5, s=6)["proof"] == good       # the handle survived every failure
    # mutated zkeys that still parse must load (or fail cleanly) on a real device too, and r1cs/ptau mutants must not
    # bring the setup kernel down
    for m in _mutations(samples["zkey"], rng, 60):
        rc = _call(lambda: api.Zkey(m).close())
        assert rc in ALLOWED | {0}
    for m in _mutations(samples["r1cs"], rng, 40):
        assert _call(lambda: api.zkey_new(m, samples["ptau"])) in ALLOWED | {0}
