"""nzcp_circom_b200 -- B200-native Groth16 (BN254) prover behind the snarkjs groth16 surface.

    from nzcp_circom_b200 import groth16
    out = groth16.prove("circuit_final.zkey", "witness.wtns")      # {"proof": ..., "publicSignals": [...]}
    ok = groth16.verify(groth16.exportVerificationKey("circuit_final.zkey"), out["publicSignals"], out["proof"])

All proving work runs in libnzcp_prover.so (hand-written sm_100a CUDA behind the C ABI in include/nzcp_prover.h).
There is no CPU fallback: importing works anywhere, proving raises NzcpError(NZCP_E_CUDA) without a GPU.
"""
from . import groth16, nzcp_input  # noqa: F401
from ._lib import NzcpError, load  # noqa: F401
from .api import Prover, SynthCircuit, Zkey  # noqa: F401
