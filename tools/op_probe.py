#!/usr/bin/env python3
"""One device-resident operation, a few times (for ncu captures and quick A/B timing):
    python tools/op_probe.py keygen|verify|verify_shared|sign|sk2pk [ps] [lanes] [reps]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fips204_b200 as F
from fips204_b200 import _lib

op = sys.argv[1] if len(sys.argv) > 1 else "keygen"
ps = int(sys.argv[2]) if len(sys.argv) > 2 else 65
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
L = _lib.lib()
m = F.PARAM_SETS[ps]
st = torch.cuda.current_stream()
sh = C.c_void_p(st.cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
g = torch.Generator(device=dev)
g.manual_seed(2)
xi = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
moff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 32
pk = torch.empty((n, m.PK_LEN), dtype=torch.uint8, device=dev)
sk = torch.empty((n, m.SK_LEN), dtype=torch.uint8, device=dev)
sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8, device=dev)
ok = torch.zeros(n, dtype=torch.uint8, device=dev)
_lib.check(L.mldsa_keygen_batch_dev(ps, sh, n, p(xi), p(pk), p(sk)))
if op == "verify":
    _lib.check(L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), m.SK_LEN, p(msg), p(moff), None, None, p(rnd), p(sig), None))
if op == "verify_shared":  # every lane under key 0
    _lib.check(L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), 0, p(msg), p(moff), None, None, p(rnd), p(sig), None))
fn = {
    "keygen": lambda: L.mldsa_keygen_batch_dev(ps, sh, n, p(xi), p(pk), p(sk)),
    "sk2pk": lambda: L.mldsa_sk_to_pk_batch_dev(ps, sh, n, p(sk), p(pk)),
    "sign": lambda: L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), m.SK_LEN, p(msg), p(moff), None, None, p(rnd), p(sig), None),
    "verify": lambda: L.mldsa_verify_batch_dev(ps, sh, 0, n, p(pk), m.PK_LEN, p(msg), p(moff), None, None, p(sig), p(ok)),
    "verify_shared": lambda: L.mldsa_verify_batch_dev(ps, sh, 0, n, p(pk), 0, p(msg), p(moff), None, None, p(sig), p(ok)),
}[op]
import os
L.mldsa_profile_enable(0 if os.environ.get("PROBE_NOPROF") else 1)
for r in range(reps):
    L.mldsa_profile_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    _lib.check(fn())
    e1.record(st)
    torch.cuda.synchronize()
    parts = []
    for i in range(L.mldsa_profile_count()):
        ms, cnt = C.c_double(), C.c_uint64()
        L.mldsa_profile_get(i, C.byref(ms), C.byref(cnt))
        if cnt.value:
            parts.append(f"{L.mldsa_profile_name(i).decode()} {ms.value:.2f}")
    if op.startswith("verify"):
        assert bool(ok.all()), "verify rejected valid signatures"
    print(f"{op}{ps} n={n}: {e0.elapsed_time(e1):.2f} ms  = {n / e0.elapsed_time(e1) / 1e3:.2f} M/s | " + "  ".join(parts), flush=True)
