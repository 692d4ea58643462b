"""ORACLE (test infrastructure, NOT product code) -- snarkjs groth16.prove restated step for step.

Follows snarkjs 0.4.12 src/groth16_prove.js (groth16Prove, buildABC1, joinABC) and ffjavascript 0.2.48
src/engine_multiexp.js (_multiExp: window table pTSizes, per-window bucket method, Horner by doubling),
src/engine_fft.js (natural-order fft/ifft) and src/engine_applykey.js (batchApplyKey: x_i *= inc^i).
Dependency pins: /root/reference/package.json:12, yarn.lock:987-999, 408-416, 1132-1135.
"parity unpinned" (SURVEY.md F1/F3): the reference has no call site, test or golden vector for this
path and the packages are not vendored; tests pin this file by (1) the Groth16 pairing equation and
(2) the toxic-waste closed form in oracle/setup.py.

Pure Python ints: use for domain sizes up to ~2^10.  Larger sizes: oracle/c (same algorithm in C).
"""
from .bn254 import R_MOD, FR_W, FR_S, MONT_R, G1, G2, ntt, to_le32
from . import formats

# ffjavascript engine_multiexp.js
PT_SIZES = [1, 1, 1, 1, 2, 3, 4, 5, 6, 7, 7, 8, 9, 10, 11, 12, 13, 13, 14, 15, 16, 16, 17, 17, 17, 17, 17, 17,
            17, 17, 17, 17]


def _log2(v):
    return v.bit_length() - 1


def multiexp(curve, bases, scalars):
    """ffjavascript _multiExp: unsigned c-bit windows over 256-bit scalars.  -> Jacobian point."""
    F = curve.F
    zero = (F.one, F.one, F.zero)
    n = len(bases)
    assert len(scalars) == n
    if n == 0:
        return zero
    c = PT_SIZES[_log2(n)]
    n_chunks = (32 * 8 - 1) // c + 1
    jb = [curve.to_jac(P) for P in bases]
    partial = []
    for ch in range(n_chunks):
        buckets = [zero] * (1 << c)
        shift = ch * c
        mask = (1 << c) - 1
        for P, s in zip(jb, scalars):
            d = (s >> shift) & mask
            if d and not F.is_zero(P[2]):
                buckets[d] = curve.jadd(buckets[d], P)
        run, acc = zero, zero
        for d in range((1 << c) - 1, 0, -1):          # sum_d d * bucket[d] by running sums
            run = curve.jadd(run, buckets[d])
            acc = curve.jadd(acc, run)
        partial.append(acc)
    res = zero
    for ch in range(n_chunks - 1, -1, -1):
        if not F.is_zero(res[2]):
            for _ in range(c):
                res = curve.jdbl(res)
        res = curve.jadd(res, partial[ch])
    return res


def build_abc1(zkey, witness):
    """groth16_prove.js buildABC1.  Values are Montgomery-form Fr ints (as the WASM buffers hold them):
    stored coef = c*R^2, witness plain  =>  montmul(coef, w) = c*w*R."""
    n = zkey["domainSize"]
    rinv = pow(MONT_R, -1, R_MOD)
    r2 = MONT_R * MONT_R % R_MOD
    out = [[0] * n, [0] * n]
    for (m, c, s, v) in zkey["coefs"]:
        stored = v * r2 % R_MOD
        out[m][c] = (out[m][c] + stored * witness[s] * rinv) % R_MOD
    cT = [a * b * rinv % R_MOD for a, b in zip(out[0], out[1])]
    return out[0], out[1], cT


def h_scalars(zkey, witness):
    """The 'H coefficients' that go into the section-9 MSM: evaluations of A*B-C on the odd coset,
    converted out of Montgomery form (joinABC + batchFromMontgomery).  No division by Z, no final iFFT."""
    n = zkey["domainSize"]
    power = _log2(n)
    aT, bT, cT = build_abc1(zkey, witness)
    inc = 25 if power == FR_S else FR_W[power + 1]            # Fr.shift = nqr^2 = 25
    odd = []
    for vec in (aT, bT, cT):
        coef = ntt(vec, inverse=True)
        k = 1
        sc = []
        for x in coef:                                        # batchApplyKey(buff, 1, inc)
            sc.append(x * k % R_MOD)
            k = k * inc % R_MOD
        odd.append(ntt(sc))
    rinv = pow(MONT_R, -1, R_MOD)
    # operands are x*R; montmul gives a*b*R; minus c*R; fromMontgomery divides by R.
    return [((a * b * rinv - c) * rinv) % R_MOD for a, b, c in zip(*odd)]


def prove(zkey, witness, r, s, return_parts=False):
    """zkey: dict from formats.read_zkey; witness: list of plain Fr ints; r, s: blinding scalars."""
    if len(witness) != zkey["nVars"]:
        raise ValueError("Invalid witness length. Circuit: %d, witness: %d" % (zkey["nVars"], len(witness)))
    npub = zkey["nPublic"]
    h = h_scalars(zkey, witness)
    A = multiexp(G1, zkey["A"], witness)
    B1 = multiexp(G1, zkey["B1"], witness)
    B2 = multiexp(G2, zkey["B2"], witness)
    C = multiexp(G1, zkey["C"], witness[npub + 1:])
    H = multiexp(G1, zkey["H"], h)
    parts = {"h": h, "A": G1.to_affine(A), "B1": G1.to_affine(B1), "B2": G2.to_affine(B2),
             "C": G1.to_affine(C), "H": G1.to_affine(H)}
    d1, d2 = G1.to_jac(zkey["vk_delta_1"]), G2.to_jac(zkey["vk_delta_2"])
    pi_a = G1.jadd(G1.jadd(A, G1.to_jac(zkey["vk_alpha_1"])), G1.jmul(d1, r))
    pi_b = G2.jadd(G2.jadd(B2, G2.to_jac(zkey["vk_beta_2"])), G2.jmul(d2, s))
    pib1 = G1.jadd(G1.jadd(B1, G1.to_jac(zkey["vk_beta_1"])), G1.jmul(d1, s))
    pi_c = G1.jadd(C, H)
    pi_c = G1.jadd(pi_c, G1.jmul(pi_a, s))
    pi_c = G1.jadd(pi_c, G1.jmul(pib1, r))
    pi_c = G1.jadd(pi_c, G1.jmul(d1, (-(r * s)) % R_MOD))
    proof = {"pi_a": G1.to_affine(pi_a), "pi_b": G2.to_affine(pi_b), "pi_c": G1.to_affine(pi_c)}
    public = witness[1:npub + 1]
    return (proof, public, parts) if return_parts else (proof, public)


def proof_to_json(proof):
    """snarkjs proof object: decimal strings of plain affine coordinates."""
    a, b, c = proof["pi_a"], proof["pi_b"], proof["pi_c"]
    return {"pi_a": [str(a[0]), str(a[1]), "1"],
            "pi_b": [[str(b[0][0]), str(b[0][1])], [str(b[1][0]), str(b[1][1])], ["1", "0"]],
            "pi_c": [str(c[0]), str(c[1]), "1"], "protocol": "groth16", "curve": "bn128"}


def proof_to_bytes(proof):
    """8 x 32-byte LE plain values A.x A.y B.x0 B.x1 B.y0 B.y1 C.x C.y (the C-ABI proof layout)."""
    a, b, c = proof["pi_a"], proof["pi_b"], proof["pi_c"]
    return b"".join(to_le32(v) for v in (a[0], a[1], b[0][0], b[0][1], b[1][0], b[1][1], c[0], c[1]))


def prove_files(zkey_bytes, wtns_bytes, r, s, return_parts=False):
    zk = formats.read_zkey(zkey_bytes)
    wt = formats.read_wtns(wtns_bytes)
    if wt["q"] != zk["r"]:
        raise ValueError("Curve of the witness does not match the curve of the proving key")
    return prove(zk, wt["witness"], r, s, return_parts)
