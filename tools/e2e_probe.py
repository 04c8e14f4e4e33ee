#!/usr/bin/env python3
"""Does a concurrent D2H / H2D copy slow the sign kernels down?  65536-lane ML-DSA-65 sign on the default stream,
alone / with a 1 GB pinned D2H / with a 1 GB pinned H2D on a side stream."""
import ctypes as C, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fips204_b200 as F
from fips204_b200 import _lib
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
L = _lib.lib(); m = F.ml_dsa_65
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g = torch.Generator(device=dev); g.manual_seed(3)
xi = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
moff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 32
pk = torch.empty((n, m.PK_LEN), dtype=torch.uint8, device=dev); sk = torch.empty((n, m.SK_LEN), dtype=torch.uint8, device=dev)
sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8, device=dev); it = torch.zeros(n, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream(); sh = C.c_void_p(st.cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
_lib.check(L.mldsa_keygen_batch_dev(65, sh, n, p(xi), p(pk), p(sk)))
big_d = torch.empty(1 << 30, dtype=torch.uint8, device=dev); big_h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
side = torch.cuda.Stream()
def sign():
    _lib.check(L.mldsa_sign_batch_dev(65, sh, 0, n, p(sk), m.SK_LEN, p(msg), p(moff), None, None, p(rnd), p(sig), p(it)))
for _ in range(3): sign()
for mode in ("alone", "d2h", "h2d", "alone", "d2h", "h2d"):
    torch.cuda.synchronize()
    if mode == "d2h":
        with torch.cuda.stream(side): big_h.copy_(big_d, non_blocking=True)
    if mode == "h2d":
        with torch.cuda.stream(side): big_d.copy_(big_h, non_blocking=True)
    t0 = time.perf_counter(); sign(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{mode:6s} sign {1e3 * (t1 - t0):7.2f} ms   (copy done after {1e3 * (t2 - t0):7.2f} ms)", flush=True)
