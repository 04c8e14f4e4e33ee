#!/usr/bin/env python3
"""End-to-end verify through the host-pointer C-ABI call (pinned buffers): distinct keys, then one shared key.
    python tools/e2e_verify.py [lanes] [reps] [ps]
Prints wall time per call, the H2D rate it implies and the device-resident time of the same batch."""
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
ok = torch.zeros(n, dtype=torch.uint8, device=dev)
sh = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
_lib.check(L.mldsa_keygen_batch_dev(ps, sh, n, p(xi), p(pk), p(sk)))
h_pk, h_msg, h_moff = pk.cpu().pin_memory(), msg.cpu().pin_memory(), moff.cpu().pin_memory()
h_sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8).pin_memory()
h_ok = torch.zeros(n, dtype=torch.uint8).pin_memory()
for shared in (False, True):
    stride = 0 if shared else m.SK_LEN
    _lib.check(L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), stride, p(msg), p(moff), None, None, p(rnd), p(sig), None))
    pstride = 0 if shared else m.PK_LEN
    for k in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(L.mldsa_verify_batch_dev(ps, sh, 0, n, p(pk), pstride, p(msg), p(moff), None, None, p(sig), p(ok)))
        torch.cuda.synchronize()
        dres = time.perf_counter() - t0
    assert bool(ok.all().item())
    h_sig.copy_(sig)
    h_sig[7, 100] ^= 1          # one corrupted lane must come back false
    torch.cuda.synchronize()
    h2d = (m.PK_LEN if shared else h_pk.numel()) + h_sig.numel() + h_msg.numel() + h_moff.numel() * 8
    best = 1e9
    for k in range(reps):
        h_ok.zero_()
        t0 = time.perf_counter()
        _lib.check(L.mldsa_verify_batch(ps, 0, 0, n, p(h_pk), pstride, p(h_msg), p(h_moff), None, None, p(h_sig), p(h_ok)))
        dt = time.perf_counter() - t0
        best = min(best, dt)
        print(f"{'shared  ' if shared else 'distinct'} e2e {1e3 * dt:8.2f} ms  {n / dt / 1e6:6.2f} M verifies/s  H2D {h2d / dt / 1e9:5.1f} GB/s", flush=True)
    good = int(h_ok.sum().item())
    print(f"{'shared  ' if shared else 'distinct'} device-resident {1e3 * dres:.2f} ms; e2e best {1e3 * best:.2f} ms = {n / best / 1e6:.2f} M/s "
          f"({h2d / best / 1e9:.1f} GB/s H2D); verdicts {good}/{n} true, lane 7 = {int(h_ok[7])}")
    assert good == n - 1 and int(h_ok[7]) == 0

import os
if os.environ.get("E2E_KEYS") == "1":
    # distinct keys again, but resident on the device (mldsa_pk_load once): only signatures and messages cross PCIe
    _lib.check(L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), m.SK_LEN, p(msg), p(moff), None, None, p(rnd), p(sig), None))
    h_sig.copy_(sig)
    h_sig[7, 100] ^= 1
    torch.cuda.synchronize()
    h = C.c_void_p()
    t0 = time.perf_counter()
    _lib.check(L.mldsa_pk_load(ps, 0, n, p(h_pk), C.byref(h)))
    dt = time.perf_counter() - t0
    print(f"mldsa_pk_load: {1e3 * dt:.2f} ms = {n / dt / 1e6:.2f} M keys/s, {L.mldsa_keyset_device_bytes(h) / 2**30:.1f} GiB resident", flush=True)
    h2d = h_sig.numel() + h_msg.numel() + h_moff.numel() * 8
    best = 1e9
    for k in range(reps):
        h_ok.zero_()
        t0 = time.perf_counter()
        _lib.check(L.mldsa_verify_batch_keys(h, 0, n, None, p(h_msg), p(h_moff), None, None, p(h_sig), p(h_ok)))
        dt = time.perf_counter() - t0
        best = min(best, dt)
        print(f"resident e2e {1e3 * dt:8.2f} ms  {n / dt / 1e6:6.2f} M verifies/s  H2D {h2d / dt / 1e9:5.1f} GB/s", flush=True)
    good = int(h_ok.sum().item())
    print(f"resident keys: e2e best {1e3 * best:.2f} ms = {n / best / 1e6:.2f} M/s; verdicts {good}/{n} true, lane 7 = {int(h_ok[7])}")
    assert good == n - 1 and int(h_ok[7]) == 0
    L.mldsa_keyset_free(h)
