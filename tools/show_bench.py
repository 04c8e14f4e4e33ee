#!/usr/bin/env python3
"""Pretty-print the last JSON line of a bench.py output file."""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head` closes the pipe early

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"{d['metric']}: {d['value']:.0f} {d['unit']}  ({d['ms_per_step']:.1f} ms/step)  e2e {d['e2e']['value']:.0f}"
      f"  e2e resident keys {(d.get('e2e_resident_keys') or {}).get('value', 0):.0f}")
for leg in (d.get("verify") or {}).get("legs", []):
    print(f"  verify {leg['keys']:9s} {leg['value']:.0f}/s device-resident  {leg.get('e2e', {}).get('value', 0):.0f}/s e2e")
if d.get("verify"):
    print("verify:", [(l["keys"], round(l["value"])) for l in d["verify"]["legs"]])
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
    ncu = v.get("ncu") or {}
    util = (f"ALU pipe {v['alu_pipe_frac'] * 100:5.1f}% of peak" if v.get("alu_pipe_frac") else
            f"ncu: fmaheavy {ncu.get('pipe_fmaheavy_pct', 0):.0f}%  L1 wavefronts {ncu.get('l1_wavefronts_pct', 0):.0f}%  issue {ncu.get('issue_active_pct', 0):.0f}%")
    print(f"  {k:12s} {v['ms_per_step']:8.2f} ms {v['share'] * 100:5.1f}%  {util}")
print("roofline:", {k: d["roofline"][k] for k in ("kernel", "achieved", "peak", "frac")} if d.get("roofline") else None)
print("sign:", d.get("sign"))
print("peaks:", d.get("int32_peaks_tops"), "cpu:", d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"))
print("check:", d.get("check"), "clocks:", d.get("clocks"), "launches:", d.get("gpu_launches"))
