#!/usr/bin/env python3
"""End-to-end keygen through the host-pointer C-ABI call (pinned buffers): python tools/e2e_keygen.py [lanes] [reps] [ps]"""
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
pk = torch.empty((n, m.PK_LEN), dtype=torch.uint8, device=dev)
sk = torch.empty((n, m.SK_LEN), dtype=torch.uint8, device=dev)
sh = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
for k in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(L.mldsa_keygen_batch_dev(ps, sh, n, p(xi), p(pk), p(sk)))
    torch.cuda.synchronize()
    dres = time.perf_counter() - t0
h_xi = xi.cpu().pin_memory()
h_pk = torch.empty((n, m.PK_LEN), dtype=torch.uint8).pin_memory()
h_sk = torch.empty((n, m.SK_LEN), dtype=torch.uint8).pin_memory()
best = 1e9
for k in range(reps):
    h_pk.zero_(); h_sk.zero_()
    t0 = time.perf_counter()
    _lib.check(L.mldsa_keygen_batch(ps, 0, n, p(h_xi), p(h_pk), p(h_sk)))
    dt = time.perf_counter() - t0
    best = min(best, dt)
    print(f"keygen{ps} e2e {1e3 * dt:8.2f} ms  {n / dt / 1e6:6.2f} M keys/s  D2H {(h_pk.numel() + h_sk.numel()) / dt / 1e9:5.1f} GB/s", flush=True)
same = bool(torch.equal(h_pk, pk.cpu())) and bool(torch.equal(h_sk, sk.cpu()))
print(f"keygen{ps} device-resident {1e3 * dres:.2f} ms; e2e best {1e3 * best:.2f} ms = {n / best / 1e6:.2f} M keys/s; bytes equal device-resident: {same}")
assert same
