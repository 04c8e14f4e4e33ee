import ctypes as C, sys
sys.path.insert(0,'.')
from fips204_b200 import _lib
import torch
torch.cuda.set_device(0)
L=_lib.lib()
names=["LOP3+SHF","IMAD+IMAD.HI","all four","IMAD","IMAD.HI","IMAD.WIDE+LOP3","SHF","LOP3","IADD3"]
for m in range(9):
    v=C.c_double(); _lib.check(L.mldsa_int32_peak_probe(m, C.byref(v)))
    print(f"mode {m} {names[m]:16s} {v.value/1e12:7.2f} Tops/s  = {v.value/148/1.965e9:6.1f} thread-ops/clk/SM @1965MHz")
