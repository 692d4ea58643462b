# NTT pipeline timing for library variants: VARIANTS="name ..." ("" = default)
for NAME in "" ${VARIANTS}; do
  if [ -n "$NAME" ]; then export NZCP_LIB_PATH=$PWD/nzcp_circom_b200/libnzcp_prover_$NAME.so; else unset NZCP_LIB_PATH; fi
  python - <<PY
import numpy as np
from nzcp_circom_b200 import api
for lg in (20,):
    n = 1 << lg
    rs = np.random.RandomState(1)
    x = rs.randint(0, 2**32, size=(3*n, 8), dtype=np.uint64).astype(np.uint32); x[:, 7] &= 0x0FFFFFFF
    api.ntt_coset(x.copy(), lg, 3)
    ms = min(api.ntt_coset(x.copy(), lg, 3) for _ in range(5))
    muls = 3 * (2 * (n // 2) * lg + n)
    print("variant=[$NAME] coset pipeline 3 x 2^%d: %.3f ms  %.1f GFrmul/s" % (lg, ms, muls / ms / 1e6))
    y = x[:n].copy(); ms1 = min(api.ntt(y.copy(), lg) for _ in range(3)); print("   forward natural 2^%d: %.3f ms" % (lg, ms1))
PY
done
