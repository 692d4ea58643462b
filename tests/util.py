"""Shared helpers for the test-suite: tiny synthetic circuits built with the ORACLE (CPU) setup."""
import random

from oracle import formats, setup
from oracle.bn254 import R_MOD

TOXIC = {"tau": 0x1234567890ABCDEF1234567, "alpha": 0xA1FA, "beta": 0xBE7A0000001, "gamma": 0x6A33A, "delta": 0xDE17A5}


def tiny_case(seed, n_constraints, n_public, n_free, toxic=TOXIC):
    """-> dict(zkey_bytes, wtns_bytes, zkey(dict), witness, constraints, n_vars, n_public)."""
    rng = random.Random(seed)
    cons, n_vars, defines = setup.random_circuit(rng, n_constraints, n_public, n_free)
    free = []
    for i in range(n_free):
        k = rng.randrange(4)
        free.append(rng.randrange(2) if k < 2 else rng.randrange(256) if k == 2 else rng.randrange(R_MOD))
    witness = setup.solve_witness(cons, n_vars, defines, n_public, free)
    zk = setup.make_zkey(cons, n_vars, n_public, toxic)
    return {"zkey_bytes": formats.write_zkey(zk), "wtns_bytes": formats.write_wtns(witness), "zkey": zk,
            "witness": witness, "constraints": cons, "n_vars": n_vars, "n_public": n_public, "toxic": toxic}


def le32(x):
    return int(x).to_bytes(32, "little")


def g1_plain_bytes(P):
    return bytes(64) if P is None else le32(P[0]) + le32(P[1])


def g2_plain_bytes(P):
    return bytes(128) if P is None else le32(P[0][0]) + le32(P[0][1]) + le32(P[1][0]) + le32(P[1][1])
