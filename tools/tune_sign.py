#!/usr/bin/env python3
"""Sweep the launch tunables (ways, kcap, acap) for the ML-DSA sign path on one GPU.
    python tools/tune_sign.py [--ps 65] [--lanes 1048576] [--configs "1,0,0 2,3,2 ..."]"""
import argparse
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fips204_b200 as F  # noqa: E402
from fips204_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ps", type=int, default=65)
ap.add_argument("--lanes", type=int, default=1 << 20)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--profile", type=int, default=0)
ap.add_argument("--configs", default="1,0,0 2,0,0 2,3,2 2,3,3 2,4,2 2,2,2 3,3,2 4,3,2")
a = ap.parse_args()

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
L = _lib.lib()
m = F.PARAM_SETS[a.ps]
n = a.lanes
g = torch.Generator(device=dev)
g.manual_seed(7)
d_xi = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
d_msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
d_rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
d_moff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 32
d_pk = torch.empty((n, m.PK_LEN), dtype=torch.uint8, device=dev)
d_sk = torch.empty((n, m.SK_LEN), dtype=torch.uint8, device=dev)
d_sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8, device=dev)
d_it = torch.zeros(n, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream()
sh = C.c_void_p(st.cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
_lib.check(L.mldsa_keygen_batch_dev(a.ps, sh, n, p(d_xi), p(d_pk), p(d_sk)))


def sign():
    _lib.check(L.mldsa_sign_batch_dev(a.ps, sh, 0, n, p(d_sk), m.SK_LEN, p(d_msg), p(d_moff), None, None, p(d_rnd),
                                      p(d_sig), p(d_it)))


ref = None
L.mldsa_profile_enable(a.profile)
import subprocess
for cfg in a.configs.split():
    w, k, ac = (int(x) for x in cfg.split(","))
    L.mldsa_set_tuning(w, k, ac)
    sign()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(a.reps):
        sign()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    chk = int(d_sig.to(torch.int64).sum().item())
    if ref is None:
        ref = chk
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    print(f"[{clk}] ways={w} kcap={k} acap={ac}: {ms:8.2f} ms  {n / ms / 1e3:7.3f} M signs/s  same_bytes={chk == ref}", flush=True)
