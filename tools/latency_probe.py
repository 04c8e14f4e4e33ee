#!/usr/bin/env python3
"""Single-operation latency through the reference-shaped API (batch calls with n = 1, host pointers): the cost of one
try_keygen / try_sign / verify when nothing is batched.  python tools/latency_probe.py [reps]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fips204_b200 as F

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for m in (F.ml_dsa_44, F.ml_dsa_65, F.ml_dsa_87):
    pk, sk = m.KG.keygen_from_seed(bytes(32))
    msg = b"latency probe message"
    sig = sk.try_sign_with_seed(bytes(32), msg, b"")
    assert pk.verify(msg, sig, b"")
    out = {}
    for name, fn in (("keygen", lambda i: m.KG.keygen_from_seed(i.to_bytes(32, "little"))),
                     ("sign", lambda i: sk.try_sign_with_seed(i.to_bytes(32, "little"), msg, b"")),
                     ("verify", lambda i: pk.verify(msg, sig, b""))):
        for i in range(20):
            fn(i)
        ts = []
        for i in range(reps):
            t0 = time.perf_counter()
            fn(i + 100)
            ts.append(time.perf_counter() - t0)
        out[name] = (1e6 * float(np.median(ts)), 1e6 * float(np.percentile(ts, 95)))
    print(f"ML-DSA-{m.PS}: " + "  ".join(f"{k} median {v[0]:7.1f} us (p95 {v[1]:7.1f})" for k, v in out.items()), flush=True)
