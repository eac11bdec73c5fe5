#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall reasons per SASS opcode class and the
hottest instructions.  Usage: ncu_stalls.py src.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); by_op = collections.defaultdict(collections.Counter)
insts = []
for r in rows[2:]:
    if r and r[0] == "Address": continue
    if len(r) < len(hdr): continue
    src = r[col["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"): op = src.split()[1]
    op = op.split(".")[0]
    n = int(r[col["# Samples"]] or 0)
    ex = int(r[col["Instructions Executed"]] or 0)
    for s in stalls:
        v = int(r[col[s]] or 0)
        tot[s] += v; by_op[op][s] += v
    insts.append((n, ex, r[col["Address"]], src))
T = sum(tot.values())
print("total samples", T)
for s, v in tot.most_common(12): print("  %-28s %8d %5.1f%%" % (s, v, 100.0 * v / T))
print("by opcode:")
for op, c in sorted(by_op.items(), key=lambda kv: -sum(kv[1].values()))[:18]:
    t = sum(c.values())
    print("  %-12s %8d %5.1f%%  %s" % (op, t, 100.0 * t / T, ", ".join("%s=%d" % (k[6:], v) for k, v in c.most_common(3))))
print("executed warp-instructions by opcode:")
exby = collections.Counter()
for n, ex, a, src in insts:
    op = src.split()[0] if src else "?"
    if op.startswith("@"): op = src.split()[1]
    exby[op.split(".")[0]] += ex
E = sum(exby.values())
for op, v in exby.most_common(24): print("  %-12s %12d %5.1f%%" % (op, v, 100.0 * v / E))
print("hottest instructions:")
for n, ex, a, src in sorted(insts, reverse=True)[:top]: print("  %7d %10d %s  %s" % (n, ex, a[-6:], src[:100]))
