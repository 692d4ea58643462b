"""XYZZ accumulate kernel (pair rounds off) on a dense 2^20 fixed-base G1 MSM, with / without the .L2::64B gather qualifier."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nzcp_circom_b200 import api  # noqa: E402

n = 1 << 20
bases = bytes(api.synth_points(77, n))
sc = np.random.RandomState(5).randint(0, 2 ** 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
sc[:, 7] &= 0x1FFFFFFF
api.tuning_set("gather_hint", int(sys.argv[1]))
with api.MsmPlan(bases, n, mode=0) as plan:
    out = None
    for _ in range(3):
        out, _ = plan.run(sc)
    print("gather_hint", sys.argv[1], "accumulate ms", round(plan.accumulate_ms(), 4), "result", out[:8].hex())
