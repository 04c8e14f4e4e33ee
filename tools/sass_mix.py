#!/usr/bin/env python3
"""Static SASS opcode mix of every kernel in the built objects (no GPU needed):
    python tools/sass_mix.py [--dir fips204_b200/build] [--filter 65] [--range kernel 0x1c10 0x6880]
Columns: FMA-heavy-pipe instructions (IMAD*, incl. the IMAD.IADD / IMAD.MOV / IMAD.SHL forms ptxas uses for plain adds,
moves and shifts), ALU-pipe instructions, LSU.  The arithmetic kernels are bound by the FMA-heavy pipe
(sm__pipe_fmaheavy_cycles_active), so adds that ptxas parks there are not free."""
import argparse
import collections
import re
import subprocess
from pathlib import Path

ap = argparse.ArgumentParser()
ap.add_argument("--dir", default=str(Path(__file__).resolve().parent.parent / "fips204_b200" / "build"))
ap.add_argument("--filter", default="")
ap.add_argument("--range", nargs=3, metavar=("KERNEL", "LO", "HI"), help="opcode histogram of one address range")
a = ap.parse_args()
line = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)")
for obj in sorted(Path(a.dir).glob("k_*.o")):
    txt = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    fn, mix = None, collections.OrderedDict()
    for ln in txt.splitlines():
        if "Function :" in ln:
            fn = subprocess.run(["c++filt", ln.split(":")[1].strip()], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn).replace("void mldsa::", "").replace("mldsa::", "")
            mix[fn] = collections.Counter()
            continue
        m = line.match(ln)
        if m and fn:
            addr, op = int(m.group(1), 16), m.group(2)
            if a.range and (a.range[0] not in fn or not (int(a.range[1], 16) <= addr < int(a.range[2], 16))):
                continue
            mix[fn][op] += 1
    for fn, c in mix.items():
        if a.filter not in fn or not c:
            continue
        if a.range:
            if a.range[0] in fn:
                print(fn)
                for op, n in c.most_common(40):
                    print(f"  {n:5d} {op}")
            continue
        tot = sum(c.values())
        g = lambda *ks: sum(v for k, v in c.items() if any(k == x or k.startswith(x + ".") for x in ks))  # noqa: E731
        imad_add = c["IMAD.IADD"] + g("IMAD.MOV") + g("IMAD.SHL") + c["IMAD.X"]
        print(f"{fn[:70]:70s} tot {tot:5d}  IMAD {c['IMAD']:4d} IMAD.HI {g('IMAD.HI'):4d} IMAD.WIDE {g('IMAD.WIDE'):4d} "
              f"IMAD-as-add/mov/shl {imad_add:4d} | IADD3 {g('IADD3'):4d} VIADD {c['VIADD']:4d} LOP3 {g('LOP3'):4d} SHF {g('SHF'):4d} "
              f"| LDS {g('LDS'):3d} STS {g('STS'):3d} LDG {g('LDG'):3d} STG {g('STG'):3d}")
