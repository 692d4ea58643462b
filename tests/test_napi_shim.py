"""The Node.js boundary without node (VERDICT r01 task 6).

js/nzcp_napi.c is compiled with -Wall -Werror against the hand-declared js/stub/node_api.h, linked with the mock N-API
runtime js/stub/napi_mock.c and libnzcp_prover.so, and DRIVEN through the mock: the addon's own exported functions run
with mock Buffers / numbers / arrays, its async work executes on a separate thread, its Promises settle.  On the CPU box
that covers argument checking and error mapping (zkeyLoad fails with the library's "no CUDA device" message); on the GPU
box the same driver produces real proofs that must equal the ctypes path byte for byte.

What this does NOT prove: behaviour under a real node (GC timing, libuv thread pool, Buffer pooling).  It does catch
what the round-1 review found by reading: silent fallbacks, synchronous proving, unbound entry points, type mistakes.
"""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JS = os.path.join(ROOT, "js")
BUILD = os.path.join(JS, "stub", "_build")
SO = os.path.join(BUILD, "nzcp_napi_mock.so")

V_BUFFER, V_PROMISE, V_ERROR, V_OBJECT, V_EXTERNAL = 4, 9, 10, 6, 5


@pytest.fixture(scope="module")
def mock(lib):
    os.makedirs(BUILD, exist_ok=True)
    libdir = os.path.join(ROOT, "nzcp_circom_b200")
    cmd = ["gcc", "-shared", "-fPIC", "-pthread", "-O1", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter",
           "-DNODE_GYP_MODULE_NAME=nzcp_napi", "-I" + os.path.join(JS, "stub"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(JS, "nzcp_napi.c"), os.path.join(JS, "stub", "napi_mock.c"),
           "-L" + libdir, "-l:libnzcp_prover.so", "-Wl,-rpath," + libdir, "-o", SO]
    subprocess.check_call(cmd)
    m = C.CDLL(SO)
    P = C.c_void_p
    for name, res, args in [
            ("mock_init", P, []), ("mock_undefined", P, []), ("mock_null", P, []), ("mock_number", P, [C.c_double]),
            ("mock_string", P, [C.c_char_p]), ("mock_buffer", P, [P, C.c_size_t]), ("mock_array", P, [C.POINTER(P), C.c_size_t]),
            ("mock_get", P, [P, C.c_char_p]), ("mock_call", P, [P, C.c_char_p, C.POINTER(P), C.c_size_t]),
            ("mock_exception_message", C.c_char_p, []), ("mock_exception_is_type_error", C.c_int, []),
            ("mock_run_pending", C.c_int, []), ("mock_kind", C.c_int, [P]), ("mock_number_value", C.c_double, [P]),
            ("mock_promise_state", C.c_int, [P]), ("mock_promise_value", P, [P]), ("mock_error_message", C.c_char_p, [P]),
            ("mock_buffer_len", C.c_size_t, [P]), ("mock_buffer_data", P, [P]), ("mock_live_refs", C.c_int, []),
            ("mock_live_works", C.c_int, []), ("mock_reset", None, [])]:
        f = getattr(m, name)
        f.restype, f.argtypes = res, args
    yield Mock(m)
    m.mock_reset()


class Mock:
    """Thin JS-like face over the driver: call(name, *args) returns a value handle or raises JsError."""

    def __init__(self, m):
        self.m = m
        self.exports = m.mock_init()
        assert self.exports, "the addon did not register itself (NAPI_MODULE)"
        self._keep = []

    def buf(self, data):
        b = (C.c_uint8 * max(1, len(data))).from_buffer_copy(bytes(data) if len(data) else b"\0")
        self._keep.append(b)
        return self.m.mock_buffer(C.addressof(b), len(data))

    def num(self, x):
        return self.m.mock_number(float(x))

    def arr(self, vals):
        a = (C.c_void_p * max(1, len(vals)))(*vals)
        return self.m.mock_array(a, len(vals))

    def call(self, name, *args):
        argv = (C.c_void_p * max(1, len(args)))(*args)
        r = self.m.mock_call(self.exports, name.encode(), argv, len(args))
        if not r:
            raise JsError(self.m.mock_exception_message().decode(), bool(self.m.mock_exception_is_type_error()))
        return r

    def settle(self, promise):
        """await: run the queued async work, then return the resolved Buffer's bytes or raise the rejection."""
        assert self.m.mock_kind(promise) == V_PROMISE
        assert self.m.mock_promise_state(promise) == 0, "the Promise settled before the worker ran: proving was synchronous"
        self.m.mock_run_pending()
        st, val = self.m.mock_promise_state(promise), self.m.mock_promise_value(promise)
        assert st in (1, 2)
        if st == 2:
            raise JsError(self.m.mock_error_message(val).decode(), False)
        n = self.m.mock_buffer_len(val)
        return C.string_at(self.m.mock_buffer_data(val), n)


class JsError(Exception):
    def __init__(self, msg, type_error):
        super().__init__(msg)
        self.type_error = type_error


def test_addon_compiles_warning_free_and_exports(mock):
    for name in ("zkeyLoad", "zkeyInfo", "prove", "proveBatch", "free"):
        assert mock.m.mock_get(mock.exports, name.encode()), name
    # syntax-only pass of the shim against the stub header alone (no mock runtime): the file is self-contained C
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-DNODE_GYP_MODULE_NAME=nzcp_napi",
                           "-I" + os.path.join(JS, "stub"), "-I" + os.path.join(ROOT, "include"), os.path.join(JS, "nzcp_napi.c")])


def test_argument_errors_are_type_errors(mock):
    with pytest.raises(JsError) as e:
        mock.call("zkeyLoad")
    assert e.value.type_error
    with pytest.raises(JsError) as e:
        mock.call("zkeyLoad", mock.num(5))
    assert e.value.type_error and "zkey must be a Buffer" in str(e.value)
    with pytest.raises(JsError) as e:
        mock.call("zkeyLoad", mock.buf(b"zkey"), mock.m.mock_string(b"gpu0"))
    assert e.value.type_error and "device" in str(e.value)
    for fn in ("prove", "proveBatch", "zkeyInfo", "free"):
        with pytest.raises(JsError) as e:
            mock.call(fn, mock.num(1), mock.buf(b"x"))
        assert e.value.type_error, fn


def test_bad_zkey_maps_to_snarkjs_style_error(mock):
    from nzcp_circom_b200 import api
    with pytest.raises(JsError) as e:
        mock.call("zkeyLoad", mock.buf(b"nope" + bytes(60)), mock.num(0))
    assert not e.value.type_error
    if api.device_count() == 0:
        assert "no CUDA device" in str(e.value) or "Invalid File format" in str(e.value)
    else:
        assert "Invalid File format" in str(e.value)
    assert mock.m.mock_live_refs() == 0 and mock.m.mock_live_works() == 0


# ------------------------------------------------------------------------------------------------ GPU: real proofs
@pytest.mark.gpu
def test_prove_through_addon_matches_ctypes_path(mock, lib):
    from nzcp_circom_b200 import api
    from util import tiny_case, le32
    c = tiny_case(seed=31, n_constraints=900, n_public=6, n_free=25)
    r, s = 0xABCDEF0123, 0x3210FEDCBA
    with api.Zkey(c["zkey_bytes"]) as zk, api.Prover(zk) as pr:
        want = pr.prove(c["wtns_bytes"], r=r, s=s)["proof"]
    h = mock.call("zkeyLoad", mock.buf(c["zkey_bytes"]), mock.num(0))
    assert mock.m.mock_kind(h) == V_EXTERNAL
    info = mock.call("zkeyInfo", h)
    assert mock.m.mock_number_value(mock.m.mock_get(info, b"nPublic")) == 6
    assert mock.m.mock_number_value(mock.m.mock_get(info, b"nVars")) == c["n_vars"]
    got = mock.settle(mock.call("prove", h, mock.buf(c["wtns_bytes"]), mock.buf(le32(r)), mock.buf(le32(s))))
    assert got == want
    # null / undefined / missing r, s = random blinding: a different, still 256-byte proof each time
    p1 = mock.settle(mock.call("prove", h, mock.buf(c["wtns_bytes"]), mock.m.mock_null(), mock.m.mock_undefined()))
    p2 = mock.settle(mock.call("prove", h, mock.buf(c["wtns_bytes"])))
    assert len(p1) == len(p2) == 256 and p1 != p2 and p1 != want
    # wrong-length or non-Buffer scalars are TypeErrors -- never a silent fallback to random blinding
    for bad in (mock.buf(bytes(31)), mock.num(7), mock.m.mock_string(b"12")):
        with pytest.raises(JsError) as e:
            mock.call("prove", h, mock.buf(c["wtns_bytes"]), bad, mock.buf(le32(s)))
        assert e.value.type_error and "32" in str(e.value)
    # snarkjs's own error texts arrive as rejections (read on the worker thread)
    with pytest.raises(JsError) as e:
        mock.settle(mock.call("prove", h, mock.buf(c["wtns_bytes"][:-32])))
    assert not e.value.type_error
    short = bytearray(c["wtns_bytes"])
    import struct
    struct.pack_into("<I", short, 12 + 12 + 36, c["n_vars"] - 1)
    with pytest.raises(JsError) as e:
        mock.settle(mock.call("prove", h, mock.buf(bytes(short))))
    assert "Invalid witness length. Circuit: %d, witness: %d" % (c["n_vars"], c["n_vars"] - 1) in str(e.value)
    # two proofs queued on ONE handle before returning to the event loop: serialised inside the addon, both correct
    pa = mock.call("prove", h, mock.buf(c["wtns_bytes"]), mock.buf(le32(r)), mock.buf(le32(s)))
    pb = mock.call("prove", h, mock.buf(c["wtns_bytes"]), mock.buf(le32(r)), mock.buf(le32(s)))
    assert mock.m.mock_promise_state(pa) == 0 and mock.m.mock_promise_state(pb) == 0
    mock.m.mock_run_pending()
    for p in (pa, pb):
        assert mock.m.mock_promise_state(p) == 1
        v = mock.m.mock_promise_value(p)
        assert C.string_at(mock.m.mock_buffer_data(v), 256) == want
    # proveBatch: nzcp_prove_batch underneath
    n = 5
    rb, sb = b"".join(le32(r + i) for i in range(n)), b"".join(le32(s + i) for i in range(n))
    out = mock.settle(mock.call("proveBatch", h, mock.arr([mock.buf(c["wtns_bytes"]) for _ in range(n)]), mock.buf(rb),
                                mock.buf(sb), mock.num(2)))
    assert len(out) == n * 256 and out[:256] == want
    with api.Zkey(c["zkey_bytes"]) as zk, api.Prover(zk) as pr:
        for i in range(n):
            assert out[256 * i:256 * (i + 1)] == pr.prove(c["wtns_bytes"], r=r + i, s=s + i)["proof"]
    with pytest.raises(JsError) as e:
        mock.call("proveBatch", h, mock.arr([mock.buf(c["wtns_bytes"])] * 2), mock.buf(le32(r)), mock.m.mock_null())
    assert e.value.type_error and "64 bytes" in str(e.value)
    # free() while a proof is in flight: the work finishes, the handle is released afterwards, later calls throw
    pc = mock.call("prove", h, mock.buf(c["wtns_bytes"]), mock.buf(le32(r)), mock.buf(le32(s)))
    mock.call("free", h)
    assert mock.settle(pc) == want
    with pytest.raises(JsError) as e:
        mock.call("prove", h, mock.buf(c["wtns_bytes"]))
    assert "freed" in str(e.value)
    assert mock.m.mock_live_refs() == 0 and mock.m.mock_live_works() == 0
