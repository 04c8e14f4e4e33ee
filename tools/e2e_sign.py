#!/usr/bin/env python3
"""End-to-end sign through the host-pointer C-ABI call (pinned buffers), without the rest of bench.py:
    python tools/e2e_sign.py [lanes] [reps] [ps]
Prints the wall time of every call and checks the bytes against the device-resident path.  Environment knobs of
the engine apply (MLDSA_PLAN, MLDSA_TAIL, MLDSA_TRACE, MLDSA_DEPTH)."""
import ctypes as C
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fips204_b200 as F
from fips204_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ps = int(sys.argv[3]) if len(sys.argv) > 3 else 65
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
L = _lib.lib()
m = F.PARAM_SETS[ps]
g = torch.Generator(device=dev)
g.manual_seed(0x204)
xi = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
moff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 32
pk = torch.empty((n, m.PK_LEN), dtype=torch.uint8, device=dev)
sk = torch.empty((n, m.SK_LEN), dtype=torch.uint8, device=dev)
sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8, device=dev)
it = torch.zeros(n, dtype=torch.int32, device=dev)
sh = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
_lib.check(L.mldsa_keygen_batch_dev(ps, sh, n, p(xi), p(pk), p(sk)))
for k in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), m.SK_LEN, p(msg), p(moff), None, None, p(rnd), p(sig), p(it)))
    torch.cuda.synchronize()
    print(f"device-resident {1e3 * (time.perf_counter() - t0):8.2f} ms", flush=True)
def prof(tag):
    out = []
    for i in range(L.mldsa_profile_count()):
        ms, cnt = C.c_double(), C.c_uint64()
        L.mldsa_profile_get(i, C.byref(ms), C.byref(cnt))
        if cnt.value:
            out.append(f"{L.mldsa_profile_name(i).decode()} {ms.value:.2f}ms/{cnt.value}")
    print(tag, "  ".join(out), flush=True)
    L.mldsa_profile_reset()
import os
PROF = os.environ.get("E2E_PROF") == "1"
if PROF:
    L.mldsa_profile_reset(); L.mldsa_profile_enable(1)
    for k in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), m.SK_LEN, p(msg), p(moff), None, None, p(rnd), p(sig), p(it)))
        print(f"device-resident, per-kernel events on {1e3 * (time.perf_counter() - t0):8.2f} ms  rounds {L.mldsa_last_sign_rounds()}", flush=True)
        prof("device-resident kernels:")
want_sig, want_it = sig.cpu(), it.cpu()
h_sk, h_msg, h_rnd, h_moff = sk.cpu().pin_memory(), msg.cpu().pin_memory(), rnd.cpu().pin_memory(), moff.cpu().pin_memory()
h_sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8).pin_memory()
h_it = torch.zeros(n, dtype=torch.int32).pin_memory()
del sig, pk, xi
best = 1e9
for k in range(reps):
    h_sig.zero_()
    t0 = time.perf_counter()
    _lib.check(L.mldsa_sign_batch(ps, 0, 0, n, p(h_sk), m.SK_LEN, p(h_msg), p(h_moff), None, None, p(h_rnd), p(h_sig), p(h_it)))
    dt = time.perf_counter() - t0
    best = min(best, dt)
    if PROF: prof("e2e kernels:")
    print(f"e2e {1e3 * dt:8.2f} ms  {n / dt / 1e6:6.3f} M signs/s  rounds {L.mldsa_last_sign_rounds()}", flush=True)
print("bytes equal device-resident:", bool(torch.equal(h_sig, want_sig)) and bool(torch.equal(h_it, want_it)),
      f"  best {1e3 * best:.2f} ms = {n / best / 1e6:.3f} M signs/s")

if os.environ.get("E2E_KEYS") == "1":
    # the same batch with the keys resident on the device (mldsa_sk_load once, then sign calls without sk traffic)
    h = C.c_void_p()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(L.mldsa_sk_load(ps, 0, n, p(h_sk), C.byref(h)))
    dt = time.perf_counter() - t0
    print(f"mldsa_sk_load: {1e3 * dt:.2f} ms = {n / dt / 1e6:.2f} M keys/s, {L.mldsa_keyset_device_bytes(h) / 2**30:.1f} GiB resident", flush=True)
    best = 1e9
    for k in range(reps):
        h_sig.zero_()
        t0 = time.perf_counter()
        _lib.check(L.mldsa_sign_batch_keys(h, 0, n, None, p(h_msg), p(h_moff), None, None, p(h_rnd), p(h_sig), p(h_it)))
        dt = time.perf_counter() - t0
        best = min(best, dt)
        print(f"e2e, resident keys {1e3 * dt:8.2f} ms  {n / dt / 1e6:6.3f} M signs/s", flush=True)
    print("bytes equal device-resident:", bool(torch.equal(h_sig, want_sig)) and bool(torch.equal(h_it, want_it)),
          f"  best {1e3 * best:.2f} ms = {n / best / 1e6:.3f} M signs/s")
    L.mldsa_keyset_free(h)
