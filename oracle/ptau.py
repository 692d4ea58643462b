"""ORACLE (test infrastructure, NOT product code) -- a PREPARED powers-of-tau file written from KNOWN toxic waste.

Restates the file snarkjs 0.4.12 produces with `powersoftau new` + contributions + `powersoftau prepare phase2`
(src/powersoftau_new.js, src/powersoftau_preparephase2.js; upstream, not vendored -- /root/reference/Makefile:31 names
powersOfTau28_hez_final_22.ptau, the real ceremony file) for the one purpose of testing the GPU `zkey new`
(nzcp_circom_b200/csrc/setup.cu): every point is a known scalar times a generator, so the key derived from this file must
equal oracle/setup.py's closed-form key.  "parity unpinned": layout from the published format, no real .ptau here.

Layout (binfileutils container "ptau", version 1; points = Montgomery affine LE, as in the .zkey):
  1  header: u32 n8 (32), q, u32 power, u32 ceremonyPower
  2  tauG1: tau^i G1, i < 2^(power+1) - 1        3  tauG2: tau^i G2, i < 2^power
  4  alphaTauG1: alpha tau^i G1, i < 2^power      5  betaTauG1: beta tau^i G1, i < 2^power       6  betaG2: beta G2
  7  contributions (u32 count = 0 here)
  12 tauG1 in Lagrange form: for p = 0 .. power+1 the 2^p points L_i^(2^p)(tau) G1 (natural order), back to back
  13 tauG2, 14 alphaTauG1, 15 betaTauG1 in Lagrange form: levels p = 0 .. power
"""
import struct

from .bn254 import (Q_MOD, R_MOD, G1, G2, G1_GEN, G2_GEN, FixedBase, to_le32, g1_to_bytes_mont, g2_to_bytes_mont)
from .formats import write_container
from .setup import lagrange_at


def write_ptau(tau, alpha, beta, power, prepared=True):
    fb1, fb2 = FixedBase(G1, G1_GEN), FixedBase(G2, G2_GEN)
    n = 1 << power
    pw = [1]
    for _ in range(2 * n - 2):
        pw.append(pw[-1] * tau % R_MOD)
    secs = [(1, struct.pack("<I", 32) + to_le32(Q_MOD) + struct.pack("<II", power, power))]
    secs.append((2, b"".join(g1_to_bytes_mont(fb1.mul(t)) for t in pw[:2 * n - 1])))
    secs.append((3, b"".join(g2_to_bytes_mont(fb2.mul(t)) for t in pw[:n])))
    secs.append((4, b"".join(g1_to_bytes_mont(fb1.mul(alpha * t % R_MOD)) for t in pw[:n])))
    secs.append((5, b"".join(g1_to_bytes_mont(fb1.mul(beta * t % R_MOD)) for t in pw[:n])))
    secs.append((6, g2_to_bytes_mont(fb2.mul(beta))))
    secs.append((7, struct.pack("<I", 0)))
    if prepared:
        lv = [lagrange_at(tau, p) for p in range(power + 2)]
        secs.append((12, b"".join(g1_to_bytes_mont(fb1.mul(x)) for p in range(power + 2) for x in lv[p])))
        secs.append((13, b"".join(g2_to_bytes_mont(fb2.mul(x)) for p in range(power + 1) for x in lv[p])))
        secs.append((14, b"".join(g1_to_bytes_mont(fb1.mul(alpha * x % R_MOD)) for p in range(power + 1) for x in lv[p])))
        secs.append((15, b"".join(g1_to_bytes_mont(fb1.mul(beta * x % R_MOD)) for p in range(power + 1) for x in lv[p])))
    return write_container(b"ptau", 1, secs)
