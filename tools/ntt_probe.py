#!/usr/bin/env python3
"""NTT / INTT microbench (BASELINE configs[1]) alone: python tools/ntt_probe.py [polys] [reps]"""
import ctypes as C
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fips204_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
L = _lib.lib()
st = torch.cuda.current_stream()
sh = C.c_void_p(st.cuda_stream)
peak = json.load(open(Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"))["hbm_gbs"]
polys = torch.randint(0, 8380417, (n, 256), dtype=torch.int32, device=dev)
for name, fn in (("ntt", L.mldsa_ntt_batch_dev), ("intt", L.mldsa_intt_batch_dev)):
    if name == "intt":
        polys.remainder_(8380417)
    for _ in range(3):
        _lib.check(fn(sh, n, C.c_void_p(polys.data_ptr())))
        if name == "ntt":
            polys.remainder_(8380417)
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        _lib.check(fn(sh, n, C.c_void_p(polys.data_ptr())))
        b.record(st)
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
        polys.remainder_(8380417)
    print(f"{name}: {best:.4f} ms  {n / best / 1e6:.3f} G polys/s  {n * 2048 / best / 1e6:.0f} GB/s = {n * 2048 / best / 1e6 / peak:.3f} of the measured copy peak")
