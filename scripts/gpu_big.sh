# one-off: a 2^21-domain proof (generic-size check beyond the NZCP shape), verified by pairing + closed form
python - <<'PY'
import sys, time
sys.path.insert(0, "tests")
import test_gpu_prove as t
from nzcp_circom_b200 import _lib
t0 = time.time()
t.test_full_size_proof_verifies_and_matches_closed_form(_lib.load(), 1100000, 100, 60000)
print("2^21 proof ok in %.1f s" % (time.time() - t0))
PY
