"""snarkjs-shaped surface: groth16.fullProve / groth16.prove / groth16.verify, zKey.exportVerificationKey.

Mirrors snarkjs 0.4.12 main.js `groth16` (upstream, not vendored: /root/reference/package.json:12,
yarn.lock:987-999).  The reference's tests build the prover input at /root/reference/test/nzcp.js:37-38
(`{ toBeSigned: bufferToBitArray(bytes), toBeSignedLen }`) and read public signals as witness[1..513] at
test/nzcp.js:41-48; `publicSignals` below is that same slice, in the same order, as decimal strings.

Same names, argument meaning and error text as snarkjs:
  prove(zkeyFileName, witnessFileName, logger=None)   file args: path | bytes-like | {"type": "mem", "data": ...}
  fullProve(input, wasmFile, zkeyFileName, logger=None)
  verify(vk_verifier, publicSignals, proof, logger=None)
Extensions (keyword-only): r=, s= inject the blinding scalars (snarkjs draws them with Fr.random() and has no hook --
bit-exact comparison needs one); device= picks the GPU.

The Python functions are synchronous (snarkjs returns Promises); the Node wrapper in js/ returns Promises.
The proving key is parsed and uploaded on first use and cached per (source, device): snarkjs re-reads the five point
sections from disk on every call.
"""
import os
import struct

from . import _lib, api, verifier
from ._lib import NzcpError

_Q = verifier.Q
_RINV_Q = pow(1 << 256, -1, _Q)

_cache = {}          # key -> (Zkey, Prover, vk-source bytes-like)


def _cache_key(zkey, device):
    if isinstance(zkey, dict) and zkey.get("type") == "mem":
        zkey = zkey["data"]
    if isinstance(zkey, str) or hasattr(zkey, "__fspath__"):
        p = os.path.abspath(os.fspath(zkey))
        st = os.stat(p)
        return ("path", p, st.st_mtime_ns, st.st_size, device)
    return ("mem", id(zkey), len(zkey), device)


def _get_prover(zkey, device):
    key = _cache_key(zkey, device)
    ent = _cache.get(key)
    if ent is None:
        data = api._as_bytes_like(zkey)
        zk = api.Zkey(data, device)
        ent = (zk, api.Prover(zk), zkey if key[0] == "mem" else None)   # keep in-memory sources alive (id() reuse)
        _cache[key] = ent
    return ent[0], ent[1]


def terminate():
    """snarkjs callers `await curve.terminate()` to stop the worker pool; here it frees the cached GPU state."""
    for zk, pr, _ in _cache.values():
        pr.close()
        zk.close()
    _cache.clear()


def _le(b, i):
    return int.from_bytes(b[32 * i:32 * i + 32], "little")


def proof_from_bytes(pb):
    """256-byte C-ABI proof -> snarkjs proof object (decimal strings of plain affine coordinates)."""
    v = [_le(pb, i) for i in range(8)]
    a_inf = v[0] == 0 and v[1] == 0
    b_inf = all(x == 0 for x in v[2:6])
    c_inf = v[6] == 0 and v[7] == 0
    return {
        "pi_a": [str(v[0]), str(v[1]), "0" if a_inf else "1"],
        "pi_b": [[str(v[2]), str(v[3])], [str(v[4]), str(v[5])], ["0", "0"] if b_inf else ["1", "0"]],
        "pi_c": [str(v[6]), str(v[7]), "0" if c_inf else "1"],
        "protocol": "groth16",
        "curve": "bn128",
    }


def _public_signals(wtns_bytes, n_public):
    """witness[1 .. nPublic] from section 2 of the .wtns image, as decimal strings."""
    mv = memoryview(wtns_bytes)
    nsec = struct.unpack_from("<I", mv, 8)[0]
    pos = 12
    for _ in range(nsec):
        sid, ln = struct.unpack_from("<IQ", mv, pos)
        pos += 12
        if sid == 2:
            return [str(int.from_bytes(mv[pos + 32 * i:pos + 32 * i + 32], "little")) for i in range(1, n_public + 1)]
        pos += ln
    raise NzcpError(_lib.NZCP_E_FORMAT, "wtns: missing section 2")


def prove(zkeyFileName, witnessFileName, logger=None, *, r=None, s=None, device=0):
    """groth16.prove -> {"proof": {...}, "publicSignals": [...]}"""
    if logger:
        logger.debug("Reading Wtns")
    wt = api._as_bytes_like(witnessFileName)
    if logger:
        logger.debug("Reading zkey")
    zk, pr = _get_prover(zkeyFileName, device)
    res = pr.prove(wt, r=r, s=s, debug=logger is not None)
    if logger:
        for k, v in res["stage_ms"].items():
            logger.debug("%s: %.3f ms" % (k, v))
    return {"proof": proof_from_bytes(res["proof"]), "publicSignals": _public_signals(wt, zk.n_public)}


def fullProve(input, wasmFile, zkeyFileName, logger=None, *, r=None, s=None, device=0):
    """groth16.fullProve: witness calculation (circom-generated WASM, CPU) followed by prove().

    The witness generator is outside the hot path and no WASM runtime ships with this package:
    `wasmFile` must be a callable `calc(input) -> .wtns bytes` (e.g. a binding of circom_runtime's
    WitnessCalculator); the Node module in js/ passes snarkjs's own `wtns.calculate` here.  A path to a .wasm raises.
    """
    if not callable(wasmFile):
        raise NzcpError(_lib.NZCP_E_ARG,
                        "fullProve: no WASM runtime in this build; pass a witness calculator callable as wasmFile "
                        "(the Node module js/index.js uses circom_runtime for this step)")
    wtns = wasmFile(input)
    return prove(zkeyFileName, {"type": "mem", "data": wtns}, logger, r=r, s=s, device=device)


def verify(vk_verifier, publicSignals, proof, logger=None):
    """groth16.verify: CPU pairing check (snarkjs verifies on the CPU as well)."""
    return verifier.verify(vk_verifier, publicSignals, proof, logger)


def exportVerificationKey(zkeyFileName):
    """snarkjs zKey.exportVerificationKey: header points + IC (section 3), plain decimal strings."""
    data = api._as_bytes_like(zkeyFileName)
    mv = memoryview(data)
    if bytes(mv[:4]) != b"zkey":
        raise NzcpError(_lib.NZCP_E_FORMAT, "zkey file: Invalid File format")
    nsec = struct.unpack_from("<I", mv, 8)[0]
    pos, secs = 12, {}
    for _ in range(nsec):
        sid, ln = struct.unpack_from("<IQ", mv, pos)
        pos += 12
        secs.setdefault(sid, (pos, ln))
        pos += ln
    if struct.unpack_from("<I", mv, secs[1][0])[0] != 1:
        raise NzcpError(_lib.NZCP_E_NOT_GROTH16, "zkey file is not groth16")
    h = secs[2][0]
    n_vars, n_public, _dom = struct.unpack_from("<III", mv, h + 72)

    def fq(off):
        return int.from_bytes(mv[off:off + 32], "little") * _RINV_Q % _Q

    def g1(off):
        x, y = fq(off), fq(off + 32)
        return [str(x), str(y), "0" if x == 0 and y == 0 else "1"]

    def g2(off):
        c = [fq(off + 32 * i) for i in range(4)]
        inf = all(v == 0 for v in c)
        return [[str(c[0]), str(c[1])], [str(c[2]), str(c[3])], ["0", "0"] if inf else ["1", "0"]]

    p = h + 84
    vk = {"protocol": "groth16", "curve": "bn128", "nPublic": n_public}
    vk["vk_alpha_1"] = g1(p)
    vk["vk_beta_2"] = g2(p + 128)
    vk["vk_gamma_2"] = g2(p + 256)
    vk["vk_delta_2"] = g2(p + 448)
    ic0 = secs[3][0]
    vk["IC"] = [g1(ic0 + 64 * i) for i in range(n_public + 1)]
    return vk
