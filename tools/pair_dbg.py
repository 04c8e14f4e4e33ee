import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import fips204_b200 as F, oracle_lib as O
from synth import synth_block
ps, n = 65, 9000
m = F.PARAM_SETS[ps]
xi = synth_block("pd-xi", n, 32)
pk, sk = m.keygen_batch(xi)
msgs = [bytes(x) for x in synth_block("pd-m", n, 32)]
rnd = synth_block("pd-r", n, 32)
sig, it = m.sign_batch(sk, msgs, rnd, want_iters=True)
print("signed", it.mean())
osig, oit = O.sign_batch(ps, sk, msgs, rnd)
print("match", (sig == osig).all(), (it == oit).all(), np.nonzero((sig != osig).any(axis=1))[0][:10])
print(m.verify_batch(pk, msgs, sig).all())
