#!/usr/bin/env python3
"""Per-kernel timing of shared-key verify, repeated, to chase run-to-run variance."""
import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fips204_b200 as F
from fips204_b200 import _lib
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
L = _lib.lib(); ps = int(sys.argv[1]) if len(sys.argv) > 1 else 65; m = F.PARAM_SETS[ps]; n = 1 << 20
g = torch.Generator(device=dev); g.manual_seed(5)
xi = torch.randint(0, 256, (1, 32), dtype=torch.uint8, device=dev, generator=g)
msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
moff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 32
pk = torch.empty((1, m.PK_LEN), dtype=torch.uint8, device=dev); sk = torch.empty((1, m.SK_LEN), dtype=torch.uint8, device=dev)
sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8, device=dev); ok = torch.zeros(n, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(); sh = C.c_void_p(st.cuda_stream); p = lambda t: C.c_void_p(t.data_ptr())
_lib.check(L.mldsa_keygen_batch_dev(ps, sh, 1, p(xi), p(pk), p(sk)))
_lib.check(L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), 0, p(msg), p(moff), None, None, p(rnd), p(sig), None))
L.mldsa_profile_enable(1)
for rep in range(8):
    L.mldsa_profile_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    _lib.check(L.mldsa_verify_batch_dev(ps, sh, 0, n, p(pk), 0, p(msg), p(moff), None, None, p(sig), p(ok)))
    e1.record(st); torch.cuda.synchronize()
    prof = {}
    for i in range(L.mldsa_profile_count()):
        ms, cnt = C.c_double(), C.c_uint64(); L.mldsa_profile_get(i, C.byref(ms), C.byref(cnt))
        if cnt.value: prof[L.mldsa_profile_name(i).decode()] = round(ms.value, 2)
    print(f"rep {rep}: {e0.elapsed_time(e1):7.2f} ms ok={bool(ok.all())}", prof, flush=True)
