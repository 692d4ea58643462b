"""The data either side of the proving path for the NZCP circuit: prover INPUT built from a pass URI, and the PUBLIC
SIGNALS the circuit must produce for it ("the callers and data formats either side of the path").

Mirrors the behaviour of the reference's test-side helpers (host logic, CPU; nothing here is on the GPU path):
  * /root/reference/test/helpers/nzcp.js:140-158  getToBeSignedAndRs: base32 payload -> COSE_Sign1 -> Sig_structure
  * /root/reference/test/nzcp.js:11-16, 37-38     prepareToBeSigned + { toBeSigned: bits, toBeSignedLen }
  * /root/reference/test/helpers/utils.js:2-10    bufferToBitArray: MSB-first bits of every byte
  * /root/reference/test/nzcp.js:18-30, 41-48     getNZCPPubIdentity and the witness[1..513] layout:
        [0..255] sha256("given,family,dob") bits, [256..511] sha256(ToBeSigned) bits, [512] exp
    (circuits/nzcptpl.circom:466-468 declares the three outputs in that order).
The reference takes givenName / familyName / dob / exp from @vaxxnz/nzcp's verifyPassURIOffline; here they are read
straight from the CWT claims of the same payload (signature verification is not part of this module).
"""
import base64
import hashlib

PREFIX = "NZCP:/1/"


class PassError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ minimal CBOR reader
def _cbor(buf, pos=0):
    """Decode one RFC 7049 item of the subset an NZCP pass uses -> (value, next position)."""
    if pos >= len(buf):
        raise PassError("truncated CBOR")
    head = buf[pos]
    major, info = head >> 5, head & 31
    pos += 1
    if info < 24:
        arg = info
    elif info in (24, 25, 26, 27):
        n = 1 << (info - 24)
        if pos + n > len(buf):
            raise PassError("truncated CBOR")
        arg = int.from_bytes(buf[pos:pos + n], "big")
        pos += n
    else:
        raise PassError("unsupported CBOR length encoding")
    if major == 0:
        return arg, pos
    if major == 1:
        return -1 - arg, pos
    if major in (2, 3):
        if pos + arg > len(buf):
            raise PassError("truncated CBOR")
        raw = bytes(buf[pos:pos + arg])
        return (raw if major == 2 else raw.decode("utf-8")), pos + arg
    if major == 4:
        out = []
        for _ in range(arg):
            v, pos = _cbor(buf, pos)
            out.append(v)
        return out, pos
    if major == 5:
        out = {}
        for _ in range(arg):
            k, pos = _cbor(buf, pos)
            v, pos = _cbor(buf, pos)
            out[k] = v
        return out, pos
    if major == 6:                       # tag: return the tagged item
        return _cbor(buf, pos)
    raise PassError("unsupported CBOR major type %d" % major)


def _bstr(data):
    n = len(data)
    if n <= 23:
        return bytes([0x40 + n]) + data
    if n < 256:
        return bytes([0x58, n]) + data
    if n < 65536:
        return bytes([0x59, n >> 8, n & 0xFF]) + data
    raise PassError("byte string too long")


def decode_pass(pass_uri):
    """-> dict(protected=bytes, payload=bytes, signature=bytes, claims=dict) of the COSE_Sign1 inside the URI."""
    if not pass_uri.startswith(PREFIX):
        raise PassError("not an NZCP:/1/ pass URI")
    b32 = pass_uri[len(PREFIX):]
    b32 += "=" * (-len(b32) % 8)
    try:
        raw = base64.b32decode(b32)
    except Exception as e:  # noqa: BLE001
        raise PassError("invalid base32 payload") from e
    if not raw or raw[0] != 0xD2:        # CBOR tag 18 = COSE_Sign1
        raise PassError("payload is not a COSE_Sign1 structure")
    item, _ = _cbor(raw, 1)
    if not (isinstance(item, list) and len(item) == 4 and isinstance(item[0], bytes) and item[1] == {}
            and isinstance(item[2], bytes) and isinstance(item[3], bytes)):
        raise PassError("malformed COSE_Sign1")
    claims, _ = _cbor(item[2], 0)
    return {"protected": item[0], "payload": item[2], "signature": item[3], "claims": claims}


def to_be_signed(pass_uri):
    """The COSE Sig_structure ["Signature1", protected, h'', payload] -- the bytes the circuit hashes and parses."""
    d = decode_pass(pass_uri)
    return b"\x84" + b"\x6aSignature1" + _bstr(d["protected"]) + _bstr(b"") + _bstr(d["payload"])


def signature_rs(pass_uri):
    sig = decode_pass(pass_uri)["signature"]
    return sig[:32].hex().upper(), sig[32:64].hex().upper()


def bits_msb_first(data):
    return [(b >> (7 - j)) & 1 for b in data for j in range(8)]


def circuit_input(pass_uri, max_len):
    """The `input` object of groth16.fullProve for NZCPPubIdentity(..., MaxToBeSignedBytes = max_len, ...):
    nzcp_exampleTest uses 314, nzcp_liveTest 355 (circuits/nzcp_exampleTest.circom:4, nzcp_liveTest.circom:4)."""
    tbs = to_be_signed(pass_uri)
    if len(tbs) > max_len:
        raise PassError("ToBeSigned is %d bytes, circuit maximum is %d" % (len(tbs), max_len))
    return {"toBeSigned": bits_msb_first(tbs + bytes(max_len - len(tbs))), "toBeSignedLen": len(tbs)}


def public_identity(pass_uri):
    """-> dict(credSubjHash, toBeSignedHash, exp) as the reference's getNZCPPubIdentity computes it."""
    d = decode_pass(pass_uri)
    claims = d["claims"]
    try:
        subj = claims["vc"]["credentialSubject"]
        concat = "%s,%s,%s" % (subj["givenName"], subj["familyName"], subj["dob"])
        exp = int(claims[4])
    except (KeyError, TypeError) as e:
        raise PassError("pass has no vc.credentialSubject / exp claim") from e
    return {"credSubjConcat": concat,
            "credSubjHash": hashlib.sha256(concat.encode("utf-8")).hexdigest(),
            "toBeSignedHash": hashlib.sha256(to_be_signed(pass_uri)).hexdigest(),
            "exp": exp}


def expected_public_signals(pass_uri):
    """The 513 publicSignals (decimal strings) a proof for this pass carries, in witness[1..513] order."""
    pid = public_identity(pass_uri)
    bits = bits_msb_first(bytes.fromhex(pid["credSubjHash"])) + bits_msb_first(bytes.fromhex(pid["toBeSignedHash"]))
    return [str(b) for b in bits] + [str(pid["exp"])]


def check_public_signals(pass_uri, public_signals):
    """True when a proof's publicSignals are the ones this pass must produce."""
    return [str(x) for x in public_signals] == expected_public_signals(pass_uri)
