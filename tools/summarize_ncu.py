#!/usr/bin/env python3
"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).

    python tools/summarize_ncu.py <tag> [--launches gpurun_out/launches_X.csv] [--rep gpurun_out/prof_X.ncu-rep]
"""
import argparse
import collections
import csv
import io
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fma%"),
    # IMAD / IMAD.HI / IMAD.WIDE all issue to the FMA-HEAVY half of the FMA pipe (an IMAD.HI holds it 4 cycles): this,
    # not pipe_fma%, is the arithmetic kernels' busiest math pipe
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu%"),
    # shared-memory + global wavefronts through the L1 data pipe (one per cycle and SM)
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_wavefronts%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("smsp__inst_executed.sum", "warp_inst"),
]


def short(name):
    n = name.replace("void ", "").replace("mldsa::", "")
    return n.split("(")[0][:44]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or r[0] == "ID" or len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "second": 1e3}.get(r[ui], 1e-6)
        k = short(r[ki])
        if k.startswith(("int_probe", "at::", "<unnamed>")):
            continue  # roofline probe and torch's input generators are not part of the step
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    out.write(f"## Launch list ({Path(path).name}: `ncu --metrics gpu__time_duration.sum --clock-control none`)\n\n")
    out.write("Per-launch times are cold-cache and serialised; compare SHARES, not absolutes.\n\n")
    out.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"| `{k}` | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |\n")
    out.write(f"| **total** | {sum(a[0] for a in agg.values())} | {tot:.3f} | 100% |\n\n")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out.write(f"## Full captures ({Path(rep).name}: `ncu --set full --clock-control none --import-source on`)\n\n")
    out.write("| kernel | " + " | ".join(n for _, n in METRICS) + " |\n|---|" + "---:|" * len(METRICS) + "\n")
    for r in data:
        cells = []
        for m, _ in METRICS:
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                    cells.append(f"{v:.4g} {units[i]}".replace(" %", "").replace("register/thread", "").strip())
                except ValueError:
                    cells.append(r[i])
            else:
                cells.append("-")
        out.write(f"| `{short(r[hdr.index('Kernel Name')])}` | " + " | ".join(cells) + " |\n")
    stall = [i for i, c in enumerate(hdr) if "pcsamp_warps_issue_stalled" in c and "not_issued" not in c]
    if stall:
        out.write("\nTop warp-stall reasons (share of PC samples):\n\n")
        for r in data:
            tot = sum(float(r[i] or 0) for i in stall) or 1.0
            top = sorted(((float(r[i] or 0) / tot, hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", "")) for i in stall), reverse=True)[:5]
            out.write(f"- `{short(r[hdr.index('Kernel Name')])}`: " + ", ".join(f"{n} {100 * v:.0f}%" for v, n in top) + "\n")
    out.write("\n")


KERNEL_KEYS = [  # kernel function name prefix -> the engine's kernel class (mldsa_profile_name), units per lane of a capture
    ("expand_a_kernel", "expand_a", 1), ("sign_key_kernel", "sign_key", 1), ("sign_prep_kernel", "sign_prep", 1),
    ("mask_kernel", "mask", 2), ("sign_w_staged_kernel", "sign_w", 2), ("ctilde_kernel", "ctilde", 2), ("sign_z_kernel", "sign_z", 2),
    ("kg_fused_kernel", "kg_arith", 1), ("kg_arith_kernel", "kg_arith_materialised", 1), ("verify_fused_kernel", "verify_arith", 1),
    ("verify_arith_kernel", "verify_arith_cached", 1), ("verify_pre_kernel", "verify_pre", 1), ("rej_bounded_kernel", "rej_bounded", 1),
    ("shake_rows_kernel", "shake_rows", 1), ("ntt_kernel<(bool)0>", "ntt", 1), ("ntt_kernel<(bool)1>", "intt", 1),
]
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def kernels_json(rep, tag, lanes, ps):
    """profiles/ncu_kernels.json: per kernel class, what ncu measured on the FIRST launch of the kernel in the capture
    (sign rounds: the first round, every lane live, two candidates per lane): pipe utilisations, DRAM bytes, duration.
    bench.py reads it for roofline.traffic and for the arithmetic kernels' pipe figures."""
    import json
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    dst = ROOT / "profiles" / "ncu_kernels.json"
    rec = json.load(open(dst)) if dst.exists() else {}

    def val(r, name, scale=False):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        return v * UNIT_SCALE.get(units[i], 1) if scale else v

    seen = set()
    for r in data:
        full_name = r[hdr.index("Kernel Name")].replace("mldsa::", "")
        for prefix, key, per_lane in KERNEL_KEYS:
            if prefix in full_name and key not in seen:
                seen.add(key)
                t_i = hdr.index("gpu__time_duration.sum")
                t_ms = float(r[t_i].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[t_i], 1e-6)
                rec[f"{key}:{ps}"] = {
                    "kernel": full_name.split("(")[0][:80], "time_ms": t_ms, "lanes": lanes, "units": lanes * per_lane,
                    "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "pipe_alu_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                    "pipe_fma_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                    "pipe_lsu_pct": val(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                    "pipe_fmaheavy_pct": val(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                    "l1_wavefronts_pct": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                    "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    "dram_read_bytes": val(r, "dram__bytes_read.sum", True), "dram_write_bytes": val(r, "dram__bytes_write.sum", True),
                    "regs": val(r, "launch__registers_per_thread"), "warp_inst": val(r, "smsp__inst_executed.sum"),
                    "source": f"profiles/{tag}.md (ncu --set full --clock-control none, {Path(rep).name}, {lanes} lanes)"}
    json.dump(rec, open(dst, "w"), indent=1, sort_keys=True)
    print(dst, sorted(seen))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--note", default="")
    ap.add_argument("--lanes", type=int, default=0, help="lanes of the captured run: also update profiles/ncu_kernels.json")
    ap.add_argument("--ps", type=int, default=65)
    a = ap.parse_args()
    dst = ROOT / "profiles" / f"{a.tag}.md"
    dst.parent.mkdir(parents=True, exist_ok=True)
    with open(dst, "w") as out:
        out.write(f"# ncu summary {a.tag}\n\n{a.note}\n\n")
        if a.launches:
            launches(a.launches, out)
        if a.rep:
            full(a.rep, out)
    print(dst)
    if a.rep and a.lanes:
        kernels_json(a.rep, a.tag, a.lanes, a.ps)
