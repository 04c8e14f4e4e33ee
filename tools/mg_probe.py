#!/usr/bin/env python3
"""Multi-rank probe (torchrun): per-call wall time of device-resident verify / sign on every rank, with and without
big pinned host allocations alive.  Chases the 2x slowdown of the verify legs seen in a 2-GPU bench run."""
import ctypes as C, os, sys, time
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fips204_b200 as F
from fips204_b200 import _lib
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if os.environ.get("MG_BIND") == "1":
    print(f"rank {rank} bound to cores {F.bind_host_to_device(local)}", flush=True)
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
L = _lib.lib(); ps = 65; m = F.PARAM_SETS[ps]; n = 1 << 20
g = torch.Generator(device=dev); g.manual_seed(5 + rank)
xi = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
moff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 32
pk = torch.empty((n, m.PK_LEN), dtype=torch.uint8, device=dev); sk = torch.empty((n, m.SK_LEN), dtype=torch.uint8, device=dev)
sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8, device=dev); ok = torch.zeros(n, dtype=torch.uint8, device=dev)
sh = C.c_void_p(torch.cuda.current_stream().cuda_stream); p = lambda t: C.c_void_p(t.data_ptr())
_lib.check(L.mldsa_keygen_batch_dev(ps, sh, n, p(xi), p(pk), p(sk)))
def sign(): _lib.check(L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), m.SK_LEN, p(msg), p(moff), None, None, p(rnd), p(sig), None))
def verify(): _lib.check(L.mldsa_verify_batch_dev(ps, sh, 0, n, p(pk), m.PK_LEN, p(msg), p(moff), None, None, p(sig), p(ok)))
def t(fn, tag, reps=4):
    out = []
    for _ in range(reps):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); out.append(1e3 * (time.perf_counter() - t0))
    print(f"rank {rank} {tag}: " + " ".join(f"{x:7.2f}" for x in out), flush=True)
sign(); t(sign, "sign  "); t(verify, "verify")
h = [torch.empty(3 << 30, dtype=torch.uint8).pin_memory() for _ in range(3)]
t(verify, "verify, 9 GB pinned alive"); t(sign, "sign, 9 GB pinned alive")
hp = pk.cpu().pin_memory(); hs = sig.cpu().pin_memory(); hm = msg.cpu().pin_memory(); hmo = moff.cpu().pin_memory(); hok = torch.zeros(n, dtype=torch.uint8).pin_memory()
def vhost(): _lib.check(L.mldsa_verify_batch(ps, local, 0, n, p(hp), m.PK_LEN, p(hm), p(hmo), None, None, p(hs), p(hok)))
t(vhost, "verify e2e"); t(verify, "verify after e2e")
hsk = sk.cpu().pin_memory(); hr = rnd.cpu().pin_memory(); hsig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8).pin_memory()
def shost(): _lib.check(L.mldsa_sign_batch(ps, local, 0, n, p(hsk), m.SK_LEN, p(hm), p(hmo), None, None, p(hr), p(hsig), None))
t(shost, "sign e2e")
if world > 1: dist.destroy_process_group()
