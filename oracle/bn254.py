"""ORACLE (test infrastructure, NOT product code) -- BN254 arithmetic in plain Python ints.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import anything under oracle/.  The product (nzcp_circom_b200/) never does.

PARITY STATUS: "parity unpinned".  The algorithm on the hot path lives in third-party packages
that are NOT vendored in /root/reference (package.json:12 -> snarkjs ^0.4.12; yarn.lock:987-999
pins snarkjs 0.4.12, yarn.lock:408-416 ffjavascript 0.2.48, yarn.lock:1132-1135 wasmcurves 0.1.0)
and the reference holds no golden vector, test or call site for the prover.
This file restates the *published* algorithms of those packages; what pins it is mathematics:
every proof must satisfy the Groth16 pairing equation and must equal the toxic-waste closed form
(oracle/setup.py).  No result here has been diffed against a live snarkjs.

What is restated, by upstream file:
  * ffjavascript src/f1field.js / wasmcurves build_f1m.js : Montgomery-form prime field, R = 2^256.
  * ffjavascript src/fft.js / engine_fft.js               : radix-2 NTT, natural order in and out,
    roots w[k] derived from the smallest quadratic non-residue (5 for Fr).
  * ffjavascript src/ec.js / wasmcurves build_curve_jacobian_a0.js : short-Weierstrass a=0 curve,
    affine + Jacobian points, (0,0) encodes infinity in affine buffers.
  * ffjavascript src/f2field.js : Fq2 = Fq[u]/(u^2+1).
"""

# ----------------------------------------------------------------------------- constants
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # Fr
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # Fq
MONT_R = 1 << 256
FR_S = 28                       # 2-adicity of r-1
FR_NQR = 5                      # smallest quadratic non-residue mod r (ffjavascript f1field.js)
FR_T = (R_MOD - 1) >> FR_S
G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
CURVE_B = 3
# twist coefficient b' = 3/(9+u)
_n = pow(82, -1, Q_MOD)
CURVE_B2 = (3 * 9 * _n % Q_MOD, (-3 * _n) % Q_MOD)

assert pow(FR_NQR, (R_MOD - 1) // 2, R_MOD) == R_MOD - 1
for _c in (2, 3, 4):
    assert pow(_c, (R_MOD - 1) // 2, R_MOD) == 1      # 5 really is the smallest non-residue

# w[k] = primitive 2^k-th root of unity; w[28] = 5^t, w[k-1] = w[k]^2 (ffjavascript f1field.js ctor)
FR_W = [0] * (FR_S + 1)
FR_W[FR_S] = pow(FR_NQR, FR_T, R_MOD)
for _k in range(FR_S, 0, -1):
    FR_W[_k - 1] = FR_W[_k] * FR_W[_k] % R_MOD
assert FR_W[0] == 1 and FR_W[1] == R_MOD - 1


# ----------------------------------------------------------------------------- byte codecs
def to_le32(x):
    return int(x).to_bytes(32, "little")


def from_le32(b):
    return int.from_bytes(b, "little")


def fq_to_mont(x):
    return x * MONT_R % Q_MOD


def fq_from_mont(x):
    return x * pow(MONT_R, -1, Q_MOD) % Q_MOD


def fr_to_mont(x):
    return x * MONT_R % R_MOD


def fr_from_mont(x):
    return x * pow(MONT_R, -1, R_MOD) % R_MOD


_RINV_Q = pow(MONT_R, -1, Q_MOD)
_RINV_R = pow(MONT_R, -1, R_MOD)


# ----------------------------------------------------------------------------- NTT (Fr)
def bitrev(i, bits):
    r_ = 0
    for _ in range(bits):
        r_ = (r_ << 1) | (i & 1)
        i >>= 1
    return r_


def ntt(a, inverse=False):
    """Natural-order radix-2 NTT of a list of plain Fr ints (ffjavascript Fr.fft / Fr.ifft).

    forward: X[k] = sum_j a[j] * w^(j k), w = FR_W[log2 n]; inverse uses w^-1 and scales by 1/n.
    """
    n = len(a)
    if n == 1:
        return list(a)
    bits = n.bit_length() - 1
    assert 1 << bits == n and bits <= FR_S
    a = [a[bitrev(i, bits)] for i in range(n)]
    for s in range(1, bits + 1):
        m = 1 << s
        wm = FR_W[s]
        if inverse:
            wm = pow(wm, -1, R_MOD)
        half = m >> 1
        tw = [1] * half
        for j in range(1, half):
            tw[j] = tw[j - 1] * wm % R_MOD
        for k in range(0, n, m):
            for j in range(half):
                t = tw[j] * a[k + j + half] % R_MOD
                u = a[k + j]
                a[k + j] = (u + t) % R_MOD
                a[k + j + half] = (u - t) % R_MOD
    if inverse:
        ninv = pow(n, -1, R_MOD)
        a = [x * ninv % R_MOD for x in a]
    return a


# ----------------------------------------------------------------------------- fields for curves
class _Fq:
    zero = 0
    one = 1

    @staticmethod
    def add(a, b):
        return (a + b) % Q_MOD

    @staticmethod
    def sub(a, b):
        return (a - b) % Q_MOD

    @staticmethod
    def mul(a, b):
        return a * b % Q_MOD

    @staticmethod
    def sqr(a):
        return a * a % Q_MOD

    @staticmethod
    def neg(a):
        return (-a) % Q_MOD

    @staticmethod
    def inv(a):
        return pow(a, -1, Q_MOD)

    @staticmethod
    def is_zero(a):
        return a == 0

    @staticmethod
    def muli(a, k):
        return a * k % Q_MOD


class _Fq2:
    zero = (0, 0)
    one = (1, 0)

    @staticmethod
    def add(a, b):
        return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)

    @staticmethod
    def sub(a, b):
        return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)

    @staticmethod
    def mul(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)

    @staticmethod
    def sqr(a):
        return ((a[0] + a[1]) * (a[0] - a[1]) % Q_MOD, 2 * a[0] * a[1] % Q_MOD)

    @staticmethod
    def neg(a):
        return ((-a[0]) % Q_MOD, (-a[1]) % Q_MOD)

    @staticmethod
    def inv(a):
        d = pow(a[0] * a[0] + a[1] * a[1], -1, Q_MOD)
        return (a[0] * d % Q_MOD, (-a[1]) * d % Q_MOD)

    @staticmethod
    def is_zero(a):
        return a[0] == 0 and a[1] == 0

    @staticmethod
    def muli(a, k):
        return (a[0] * k % Q_MOD, a[1] * k % Q_MOD)


class Curve:
    """y^2 = x^3 + b over field F.  Points: affine (x, y) or None; Jacobian (X, Y, Z) with Z=0 <=> inf."""

    def __init__(self, F, b, gen):
        self.F = F
        self.b = b
        self.gen = gen

    def on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.sqr(y) == F.add(F.mul(F.sqr(x), x), self.b)

    def to_jac(self, P):
        return (self.F.one, self.F.one, self.F.zero) if P is None else (P[0], P[1], self.F.one)

    def to_affine(self, J):
        F = self.F
        if F.is_zero(J[2]):
            return None
        zi = F.inv(J[2])
        zi2 = F.sqr(zi)
        return (F.mul(J[0], zi2), F.mul(J[1], F.mul(zi2, zi)))

    def jdbl(self, P):
        F = self.F
        X, Y, Z = P
        if F.is_zero(Z):
            return P
        A = F.sqr(X)
        B = F.sqr(Y)
        C = F.sqr(B)
        D = F.muli(F.sub(F.sub(F.sqr(F.add(X, B)), A), C), 2)
        E = F.muli(A, 3)
        X3 = F.sub(F.sqr(E), F.muli(D, 2))
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.muli(C, 8))
        Z3 = F.muli(F.mul(Y, Z), 2)
        return (X3, Y3, Z3)

    def jadd(self, P, Q):
        F = self.F
        if F.is_zero(P[2]):
            return Q
        if F.is_zero(Q[2]):
            return P
        Z1Z1 = F.sqr(P[2])
        Z2Z2 = F.sqr(Q[2])
        U1 = F.mul(P[0], Z2Z2)
        U2 = F.mul(Q[0], Z1Z1)
        S1 = F.mul(P[1], F.mul(Q[2], Z2Z2))
        S2 = F.mul(Q[1], F.mul(P[2], Z1Z1))
        if U1 == U2:
            if S1 == S2:
                return self.jdbl(P)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1)
        Rr = F.sub(S2, S1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.muli(V, 2))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(P[2], Q[2]), H)
        return (X3, Y3, Z3)

    def jneg(self, P):
        return (P[0], self.F.neg(P[1]), P[2])

    def jmul(self, P, k):
        k = int(k)
        if k < 0:
            return self.jmul(self.jneg(P), -k)
        acc = (self.F.one, self.F.one, self.F.zero)
        for bit in bin(k)[2:] if k else "":
            acc = self.jdbl(acc)
            if bit == "1":
                acc = self.jadd(acc, P)
        return acc

    # affine conveniences
    def add(self, P, Q):
        return self.to_affine(self.jadd(self.to_jac(P), self.to_jac(Q)))

    def mul(self, P, k):
        return self.to_affine(self.jmul(self.to_jac(P), k))

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))


G1 = Curve(_Fq, CURVE_B, G1_GEN)
G2 = Curve(_Fq2, CURVE_B2, G2_GEN)
assert G1.on_curve(G1_GEN) and G2.on_curve(G2_GEN)


class FixedBase:
    """Windowed fixed-base multiplication k*G (4-bit windows) -- used by the synthetic setup only."""

    def __init__(self, curve, base, bits=254, w=4):
        self.curve, self.w = curve, w
        self.tables = []
        P = curve.to_jac(base)
        for _ in range((bits + w - 1) // w):
            row = [None]
            acc = (curve.F.one, curve.F.one, curve.F.zero)
            for _j in range(1, 1 << w):
                acc = curve.jadd(acc, P)
                row.append(acc)
            self.tables.append(row)
            for _j in range(w):
                P = curve.jdbl(P)

    def mul(self, k):
        c = self.curve
        acc = (c.F.one, c.F.one, c.F.zero)
        i = 0
        mask = (1 << self.w) - 1
        while k:
            d = k & mask
            if d:
                acc = c.jadd(acc, self.tables[i][d])
            k >>= self.w
            i += 1
        return c.to_affine(acc)


# ----------------------------------------------------------------------------- point codecs
# zkey point encoding (snarkjs zkey_utils.js / ffjavascript ec.js toRprLEM): affine coordinates,
# 32-byte little-endian, *Montgomery form*; infinity = all-zero bytes.
def g1_to_bytes_mont(P):
    if P is None:
        return bytes(64)
    return to_le32(fq_to_mont(P[0])) + to_le32(fq_to_mont(P[1]))


def g1_from_bytes_mont(b):
    if b == bytes(64):
        return None
    return (from_le32(b[:32]) * _RINV_Q % Q_MOD, from_le32(b[32:64]) * _RINV_Q % Q_MOD)


def g2_to_bytes_mont(P):
    if P is None:
        return bytes(128)
    return b"".join(to_le32(fq_to_mont(c)) for c in (P[0][0], P[0][1], P[1][0], P[1][1]))


def g2_from_bytes_mont(b):
    if b == bytes(128):
        return None
    v = [from_le32(b[i * 32:(i + 1) * 32]) * _RINV_Q % Q_MOD for i in range(4)]
    return ((v[0], v[1]), (v[2], v[3]))
