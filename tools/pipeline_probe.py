#!/usr/bin/env python3
"""BASELINE config 5 through mldsa_pipeline_batch[_multi]: keygen -> sign -> verify chained on the device(s).
    python tools/pipeline_probe.py [lanes] [reps] [ps] [ndev]
Pinned host buffers; prints the wall time of every call, the per-stage device times and ops/s; checks a sample of
lanes against the three separate calls."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fips204_b200 as F
from fips204_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ps = int(sys.argv[3]) if len(sys.argv) > 3 else 87
ndev = int(sys.argv[4]) if len(sys.argv) > 4 else 1
L = _lib.lib()
m = F.PARAM_SETS[ps]


def pinned(nbytes, dtype=np.uint8):
    p = L.mldsa_host_alloc(nbytes)
    assert p
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dtype)


rng = np.random.default_rng(5)
xi = pinned(n * 32)
xi[:] = rng.integers(0, 256, n * 32, dtype=np.uint8)
msg = pinned(n * 32)
msg[:] = rng.integers(0, 256, n * 32, dtype=np.uint8)
rnd = pinned(n * 32)
rnd[:] = rng.integers(0, 256, n * 32, dtype=np.uint8)
moff = pinned((n + 1) * 8, np.uint64)
moff[:] = np.arange(n + 1, dtype=np.uint64) * 32
pk = pinned(n * m.PK_LEN)
sig = pinned(n * m.SIG_LEN)
ok = pinned(n)
devs = (C.c_int * ndev)(*range(ndev))
ms = (C.c_double * 3)()
for r in range(reps):
    t0 = time.perf_counter()
    rc = L.mldsa_pipeline_batch_multi(ps, devs, ndev, 0, n, xi.ctypes.data, msg.ctypes.data, moff.ctypes.data, None, None,
                                      rnd.ctypes.data, pk.ctypes.data, None, sig.ctypes.data, ok.ctypes.data, None, ms)
    dt = time.perf_counter() - t0
    _lib.check(rc)
    print(f"pipeline{ps} n={n} ndev={ndev}: {1e3 * dt:8.2f} ms  = {n / dt / 1e6:6.3f} M lanes/s ({3 * n / dt / 1e6:6.3f} M ops/s); "
          f"stages keygen {ms[0]:.2f} sign {ms[1]:.2f} verify {ms[2]:.2f} ms; all ok: {bool(ok.all())}", flush=True)
k = min(n, 2000)
pk2, sk2 = m.keygen_batch(xi[:k * 32])
sig2 = m.sign_batch(sk2, [bytes(msg[32 * i:32 * i + 32]) for i in range(k)], rnd[:k * 32])
print("sample equals three separate calls:", bool((pk2.ravel() == pk[:k * m.PK_LEN]).all() and (sig2.ravel() == sig[:k * m.SIG_LEN]).all()))
