#!/usr/bin/env python3
"""One first-round sign call (or one keygen) over n lanes of ONE input class, for an ncu instruction / memory-transaction
count (`tools/ct_trace.sh` runs it for two classes and diffs the counts): if the kernels that touch secret data execute
the same number of instructions, shared-memory wavefronts and L2 sectors for every class, what is left of a timing
difference between classes is not in the program.
    python tools/ct_trace.py sign|keygen A|B|R [ps] [lanes]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fips204_b200 as F
from fips204_b200 import _lib

op, cls = sys.argv[1], sys.argv[2]
ps = int(sys.argv[3]) if len(sys.argv) > 3 else 65
n = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
L = _lib.lib()
m = F.PARAM_SETS[ps]
sh = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
g = torch.Generator(device=dev)
g.manual_seed(0xD0DEC7)
xi = {"A": torch.full((n, 32), 0xAA, dtype=torch.uint8, device=dev), "B": torch.full((n, 32), 0x55, dtype=torch.uint8, device=dev),
      "R": torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)}[cls]
pk = torch.empty((n, m.PK_LEN), dtype=torch.uint8, device=dev)
sk = torch.empty((n, m.SK_LEN), dtype=torch.uint8, device=dev)
_lib.check(L.mldsa_keygen_batch_dev(ps, sh, n, p(xi), p(pk), p(sk)))
if op == "sign":
    g.manual_seed(1)
    msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    moff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 32
    sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8, device=dev)
    L.mldsa_debug_set_max_rounds(1)
    rc = L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), m.SK_LEN, p(msg), p(moff), None, None, p(rnd), p(sig), None)
    assert rc in (0, -5)
torch.cuda.synchronize()
