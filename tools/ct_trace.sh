#!/bin/bash
# tools/ct_trace.sh [ps] -> gpurun_out/ct_trace_<ps>.txt: executed instructions, shared-memory wavefronts and L2 sectors of every
# kernel of a first-round sign call and of a keygen, for input classes A (every lane the key of xi = 0xAA..) and R (random keys).
cd "$(dirname "$0")/.."
PS=${1:-65}
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_sectors.sum
mkdir -p gpurun_out
OUT=gpurun_out/ct_trace_${PS}.txt
: > $OUT
for op in sign keygen; do
  for cls in A R; do
    ncu --metrics $M --csv --log-file gpurun_out/ct_trace_${op}_${cls}.csv python tools/ct_trace.py $op $cls $PS > /dev/null 2>&1
  done
  python - $op $PS >> $OUT <<'PY'
import csv, sys
op, ps = sys.argv[1], sys.argv[2]
def load(cls):
    rows = [r for r in csv.reader(open(f"gpurun_out/ct_trace_{op}_{cls}.csv")) if len(r) > 10]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    out, seen = {}, {}
    for r in rows:
        if r is hdr or r[0] == "ID": continue
        name = r[ki].replace("mldsa::", "").split("(")[0].replace("void ", "")
        k = seen.setdefault((r[0], name), len([x for x in seen if x[1] == name]))
        out[(name, k, r[mi])] = r[vi]
    return out
a, r = load("A"), load("R")
own = ("keygen" if op == "keygen" else "sign", "kg_", "rej_bounded", "shake_rows", "mask", "ctilde", "sign_", "expand_a", "compact")
print(f"## ML-DSA-{ps} {op}: class A (all lanes xi = 0xAA..) against class R (random keys), 16384 lanes, per kernel launch")
print("| kernel (launch #) | metric | class A | class R | equal |")
print("|---|---|---:|---:|---|")
for key in sorted(a):
    name, k, metric = key
    if op == "sign" and (name.startswith("kg_") or name.startswith("rej_bounded") or "shake_rows" in name or "at::" in name): continue
    if "at::" in name or "elementwise" in name or "distribution" in name: continue
    if op == "keygen" and k == 0 and False: continue
    va, vr = a[key], r.get(key, "-")
    print(f"| `{name[:48]}` ({k}) | {metric.replace('.sum','')} | {va} | {vr} | {'yes' if va == vr else '**no**'} |")
PY
done
cat $OUT | grep -c "yes"; grep "no\*\*" $OUT | head -40
