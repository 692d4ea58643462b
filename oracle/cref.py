"""ORACLE (test infrastructure, NOT product code) -- ctypes binding of oracle/c (the C restatement of snarkjs
groth16.prove for the host CPU).  Checker for sizes the pure-Python oracle cannot reach, and the reported CPU baseline
(bench.py cpu_baseline / --impl reference).  "parity unpinned": see oracle/c/nzcp_oracle.c.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libnzcp_oracle.so")
STAGES = ("parse", "buildABC1", "fft_join", "msm_a", "msm_b1", "msm_b2", "msm_c", "msm_h")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")])
        lib = C.CDLL(LIB_PATH)
        lib.nzo_max_threads.restype = C.c_int
        lib.nzo_field_op.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t]
        lib.nzo_ntt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.nzo_ntt_coset.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.nzo_msm.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
        lib.nzo_prove.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        _lib = lib
    return _lib


def _addr(buf):
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p).value
    if hasattr(buf, "ctypes"):
        return buf.ctypes.data
    mv = memoryview(buf)
    return C.addressof((C.c_char * mv.nbytes).from_buffer(mv))


def max_threads():
    """Host threads available to this process (not OMP_NUM_THREADS: torchrun pins that to 1)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def field_op(field, op, a, b, n):
    out = bytearray(32 * n)
    load().nzo_field_op(field, op, bytes(a), bytes(b), _addr(out), n)
    return bytes(out)


def ntt(data, log_n, inverse=False, threads=1):
    assert load().nzo_ntt(_addr(data), log_n, int(inverse), threads) == 0


def ntt_coset(data, log_n, threads=1):
    assert load().nzo_ntt_coset(_addr(data), log_n, threads) == 0


def msm(bases, scalars, n, g2=False, threads=1):
    out = bytearray(128 if g2 else 64)
    assert load().nzo_msm(_addr(bases), _addr(scalars), n, int(g2), _addr(out), threads) == 0
    return bytes(out)


def prove(zkey, wtns, r, s, threads=0, want_h_size=0):
    """-> dict(proof=256 B, msm_a, msm_b1, msm_b2, msm_c, msm_h, [h], stage_sec).  threads=0: all host threads."""
    lib = load()
    if threads <= 0:
        threads = max_threads()
    proof = bytearray(256)
    parts = bytearray(384)
    hbuf = bytearray(want_h_size * 32) if want_h_size else None
    st = (C.c_double * 8)()
    rc = lib.nzo_prove(_addr(zkey), len(zkey), _addr(wtns), len(wtns), int(r).to_bytes(32, "little"),
                       int(s).to_bytes(32, "little"), _addr(proof), _addr(parts), _addr(hbuf) if hbuf is not None else None,
                       threads, st)
    if rc:
        raise ValueError({-2: "Invalid File format", -3: "zkey file is not groth16",
                          -4: "Curve of the witness does not match the curve of the proving key",
                          -5: "Invalid witness length"}.get(rc, "error %d" % rc))
    out = {"proof": bytes(proof), "msm_a": bytes(parts[:64]), "msm_b1": bytes(parts[64:128]),
           "msm_b2": bytes(parts[128:256]), "msm_c": bytes(parts[256:320]), "msm_h": bytes(parts[320:384]),
           "stage_sec": dict(zip(STAGES, list(st))), "threads": threads}
    if hbuf is not None:
        out["h"] = bytes(hbuf)
    return out
