#!/usr/bin/env python3
"""Does clock sampling perturb the timed region?  Device-resident 2^20-lane ML-DSA-65 sign, wall time per call:
no sampler / `nvidia-smi -lms 200` subprocess / in-process NVML thread at 100 ms."""
import ctypes as C, subprocess, sys, threading, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fips204_b200 as F
from fips204_b200 import _lib
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
L = _lib.lib(); ps = 65; m = F.PARAM_SETS[ps]; n = 1 << 20
g = torch.Generator(device=dev); g.manual_seed(5)
xi = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
moff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 32
pk = torch.empty((n, m.PK_LEN), dtype=torch.uint8, device=dev); sk = torch.empty((n, m.SK_LEN), dtype=torch.uint8, device=dev)
sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8, device=dev)
sh = C.c_void_p(torch.cuda.current_stream().cuda_stream); p = lambda t: C.c_void_p(t.data_ptr())
_lib.check(L.mldsa_keygen_batch_dev(ps, sh, n, p(xi), p(pk), p(sk)))
def sign(): _lib.check(L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), m.SK_LEN, p(msg), p(moff), None, None, p(rnd), p(sig), None))
def run(tag):
    out = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); sign(); out.append(1e3 * (time.perf_counter() - t0))
    print(f"{tag:28s}" + " ".join(f"{x:7.2f}" for x in out), flush=True)
for _ in range(2): sign()
run("no sampler")
q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def with_thread(tag, fn, period):
    global stop
    stop = False
    cnt = [0]
    def pump():
        while not stop:
            fn(); cnt[0] += 1
            time.sleep(period)
    th = threading.Thread(target=pump); th.start(); time.sleep(0.3)
    out = []
    for _ in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter(); sign(); out.append(1e3 * (time.perf_counter() - t0))
    stop = True; th.join()
    print(f"{tag:34s} mean {sum(out)/len(out):7.2f}  max {max(out):7.2f}  ({cnt[0]} queries)", flush=True)
with_thread("no query (sleep only)", lambda: None, 0.1)
with_thread("clock SM @100ms", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), 0.1)
with_thread("event reasons @100ms", lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h), 0.1)
with_thread("power @100ms", lambda: pynvml.nvmlDeviceGetPowerUsage(h), 0.1)
with_thread("clock+reasons @400ms", lambda: (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)), 0.4)
pr = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "400", "-i", "0"], stdout=subprocess.DEVNULL)
time.sleep(0.5)
with_thread("nvidia-smi -lms 400 (9 fields)", lambda: None, 0.1)
pr.terminate(); pr.wait()
