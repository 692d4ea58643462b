"""ORACLE (test infrastructure, NOT product code) -- iden3 binary containers: .zkey, .wtns, .r1cs.

Restates @iden3/binfileutils 0.0.10 (readBinFile / startWriteSection, yarn.lock:385-388 region) and
snarkjs 0.4.12 src/zkey_utils.js (readHeader, writeHeader), src/wtns_utils.js, r1csfile.
"parity unpinned": none of these packages is vendored in /root/reference; the layout
follows SURVEY.md section 8b.  All integers little-endian.

Container: 4-byte magic, u32 version, u32 nSections, then per section: u32 id, u64 byteLen, payload.
"""
import struct

from .bn254 import (R_MOD, Q_MOD, to_le32, from_le32, g1_to_bytes_mont, g1_from_bytes_mont,
                    g2_to_bytes_mont, g2_from_bytes_mont, MONT_R)


def write_container(magic, version, sections):
    """sections: list of (id, bytes)."""
    out = [magic, struct.pack("<II", version, len(sections))]
    for sid, payload in sections:
        out.append(struct.pack("<IQ", sid, len(payload)))
        out.append(payload)
    return b"".join(out)


def read_container(buf, magic):
    if bytes(buf[:4]) != magic:
        raise ValueError("%s file: bad magic" % magic.decode())
    version, nsec = struct.unpack_from("<II", buf, 4)
    pos = 12
    sections = {}
    for _ in range(nsec):
        sid, ln = struct.unpack_from("<IQ", buf, pos)
        pos += 12
        sections.setdefault(sid, []).append((pos, ln))
        pos += ln
    return version, sections


def _sec(buf, sections, sid):
    pos, ln = sections[sid][0]
    return memoryview(buf)[pos:pos + ln]


# ------------------------------------------------------------------ wtns  (snarkjs wtns_utils.js)
def write_wtns(witness):
    hdr = struct.pack("<I", 32) + to_le32(R_MOD) + struct.pack("<I", len(witness))
    body = b"".join(to_le32(w) for w in witness)
    return write_container(b"wtns", 2, [(1, hdr), (2, body)])


def read_wtns(buf):
    _, secs = read_container(buf, b"wtns")
    h = _sec(buf, secs, 1)
    n8 = struct.unpack_from("<I", h, 0)[0]
    qv = int.from_bytes(h[4:4 + n8], "little")
    nw = struct.unpack_from("<I", h, 4 + n8)[0]
    body = _sec(buf, secs, 2)
    wit = [int.from_bytes(body[i * n8:(i + 1) * n8], "little") for i in range(nw)]
    return {"n8": n8, "q": qv, "nWitness": nw, "witness": wit}


# ------------------------------------------------------------------ zkey  (snarkjs zkey_utils.js)
def write_zkey(zk):
    """zk: dict with nVars,nPublic,domainSize, vk_* points (plain affine), IC, coefs[(m,c,s,val)],
    A,B1,B2,C,H point lists (plain affine / None)."""
    s1 = struct.pack("<I", 1)  # groth16
    s2 = (struct.pack("<I", 32) + to_le32(Q_MOD) + struct.pack("<I", 32) + to_le32(R_MOD)
          + struct.pack("<III", zk["nVars"], zk["nPublic"], zk["domainSize"])
          + g1_to_bytes_mont(zk["vk_alpha_1"]) + g1_to_bytes_mont(zk["vk_beta_1"])
          + g2_to_bytes_mont(zk["vk_beta_2"]) + g2_to_bytes_mont(zk["vk_gamma_2"])
          + g1_to_bytes_mont(zk["vk_delta_1"]) + g2_to_bytes_mont(zk["vk_delta_2"]))
    s3 = b"".join(g1_to_bytes_mont(P) for P in zk["IC"])
    # coefficient values are stored as coef * R^2 mod r  (snarkjs zkey_new.js, "R2r"), SURVEY F7
    r2 = MONT_R * MONT_R % R_MOD
    s4 = [struct.pack("<I", len(zk["coefs"]))]
    for (m, c, s, v) in zk["coefs"]:
        s4.append(struct.pack("<III", m, c, s) + to_le32(v * r2 % R_MOD))
    s4 = b"".join(s4)
    s5 = b"".join(g1_to_bytes_mont(P) for P in zk["A"])
    s6 = b"".join(g1_to_bytes_mont(P) for P in zk["B1"])
    s7 = b"".join(g2_to_bytes_mont(P) for P in zk["B2"])
    s8 = b"".join(g1_to_bytes_mont(P) for P in zk["C"])
    s9 = b"".join(g1_to_bytes_mont(P) for P in zk["H"])
    s10 = to_le32(0) * 2 + struct.pack("<I", 0)  # csHash (64 B) + nContributions = 0
    return write_container(b"zkey", 1, [(1, s1), (2, s2), (3, s3), (4, s4), (5, s5), (6, s6),
                                        (7, s7), (8, s8), (9, s9), (10, s10)])


def read_zkey_header(buf):
    _, secs = read_container(buf, b"zkey")
    if struct.unpack_from("<I", _sec(buf, secs, 1), 0)[0] != 1:
        raise ValueError("zkey file is not groth16")
    h = _sec(buf, secs, 2)
    p = 0
    n8q = struct.unpack_from("<I", h, p)[0]; p += 4
    qv = int.from_bytes(h[p:p + n8q], "little"); p += n8q
    n8r = struct.unpack_from("<I", h, p)[0]; p += 4
    rv = int.from_bytes(h[p:p + n8r], "little"); p += n8r
    nVars, nPublic, domainSize = struct.unpack_from("<III", h, p); p += 12
    z = {"n8q": n8q, "q": qv, "n8r": n8r, "r": rv, "nVars": nVars, "nPublic": nPublic,
         "domainSize": domainSize, "power": domainSize.bit_length() - 1}
    z["vk_alpha_1"] = g1_from_bytes_mont(bytes(h[p:p + 64])); p += 64
    z["vk_beta_1"] = g1_from_bytes_mont(bytes(h[p:p + 64])); p += 64
    z["vk_beta_2"] = g2_from_bytes_mont(bytes(h[p:p + 128])); p += 128
    z["vk_gamma_2"] = g2_from_bytes_mont(bytes(h[p:p + 128])); p += 128
    z["vk_delta_1"] = g1_from_bytes_mont(bytes(h[p:p + 64])); p += 64
    z["vk_delta_2"] = g2_from_bytes_mont(bytes(h[p:p + 128])); p += 128
    return z, secs


def read_zkey(buf):
    """Full decode into Python objects -- small files only."""
    z, secs = read_zkey_header(buf)

    def g1s(sid):
        b = _sec(buf, secs, sid)
        return [g1_from_bytes_mont(bytes(b[i:i + 64])) for i in range(0, len(b), 64)]

    def g2s(sid):
        b = _sec(buf, secs, sid)
        return [g2_from_bytes_mont(bytes(b[i:i + 128])) for i in range(0, len(b), 128)]

    z["IC"] = g1s(3)
    c = _sec(buf, secs, 4)
    ncoef = struct.unpack_from("<I", c, 0)[0]
    r2inv = pow(MONT_R * MONT_R, -1, R_MOD)
    coefs = []
    for i in range(ncoef):
        o = 4 + i * 44
        m, cc, s = struct.unpack_from("<III", c, o)
        coefs.append((m, cc, s, from_le32(c[o + 12:o + 44]) * r2inv % R_MOD))
    z["coefs"] = coefs
    z["coefs_raw_section"] = bytes(c)
    z["A"], z["B1"], z["B2"], z["C"], z["H"] = g1s(5), g1s(6), g2s(7), g1s(8), g1s(9)
    return z


# ------------------------------------------------------------------ r1cs  (iden3 r1csfile)
def write_r1cs(n_wires, n_pub_out, n_pub_in, n_prv_in, constraints):
    """constraints: list of (A, B, C), each a dict {wire: coef} (plain Fr ints)."""
    hdr = (struct.pack("<I", 32) + to_le32(R_MOD)
           + struct.pack("<IIIIQI", n_wires, n_pub_out, n_pub_in, n_prv_in, n_wires, len(constraints)))
    body = []
    for lcs in constraints:
        for lc in lcs:
            body.append(struct.pack("<I", len(lc)))
            for wire in sorted(lc):
                body.append(struct.pack("<I", wire) + to_le32(lc[wire] % R_MOD))
    labels = b"".join(struct.pack("<Q", i) for i in range(n_wires))
    return write_container(b"r1cs", 1, [(1, hdr), (2, b"".join(body)), (3, labels)])


def read_r1cs(buf):
    _, secs = read_container(buf, b"r1cs")
    h = _sec(buf, secs, 1)
    n8 = struct.unpack_from("<I", h, 0)[0]
    prime = int.from_bytes(h[4:4 + n8], "little")
    n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_cons = struct.unpack_from("<IIIIQI", h, 4 + n8)
    b = _sec(buf, secs, 2)
    p = 0
    cons = []
    for _ in range(n_cons):
        lcs = []
        for _k in range(3):
            cnt = struct.unpack_from("<I", b, p)[0]; p += 4
            lc = {}
            for _t in range(cnt):
                wire = struct.unpack_from("<I", b, p)[0]; p += 4
                lc[wire] = int.from_bytes(b[p:p + n8], "little"); p += n8
            lcs.append(lc)
        cons.append(tuple(lcs))
    return {"prime": prime, "nVars": n_wires, "nPubOut": n_pub_out, "nPubIn": n_pub_in,
            "nPrvIn": n_prv_in, "nPublic": n_pub_out + n_pub_in, "nConstraints": n_cons,
            "constraints": cons}
