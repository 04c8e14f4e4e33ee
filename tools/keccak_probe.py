#!/usr/bin/env python3
"""Measure Keccak-f[1600] throughput for each ALU/FMA rotation split compiled into k_probe.cu."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from fips204_b200 import _lib  # noqa: E402

torch.cuda.set_device(0)
L = _lib.lib()
v = 0
while True:
    r, m = C.c_double(), C.c_uint32()
    if L.mldsa_keccak_probe(v, C.byref(r), C.byref(m)) != 0:
        break
    print(f"variant {v}: FMA sites {bin(m.value).count('1'):2d}/29 mask {m.value:#010x}  {r.value / 1e9:6.3f} G perm/s "
          f"= {r.value * 4320 / 1e12:6.2f} Tops/s algorithmic")
    v += 1
