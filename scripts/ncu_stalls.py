#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export (SASS view): total samples by stall reason and the
top-N instructions by sample count."""
import csv
import sys
from collections import Counter

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = Counter()
insts = []
opc = Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        ns = int(r[col["# Samples"]])
    except ValueError:
        continue
    for s in stall_cols:
        try:
            tot[s] += int(r[col[s]])
        except ValueError:
            pass
    insts.append((ns, r[col["Source"]].strip(), {s: int(r[col[s]] or 0) for s in stall_cols if (r[col[s]] or "0") != "0"},
                  int(r[col["Instructions Executed"]] or 0)))
    opc[r[col["Source"]].split()[0] if r[col["Source"]].split()[0][0] != "@" else r[col["Source"]].split()[1]] += int(r[col["Instructions Executed"]] or 0)
allsamp = sum(tot.values())
print("stall totals:")
for s, n in tot.most_common():
    if n:
        print(f"  {s:28s} {n:8d} {100.0 * n / allsamp:5.1f}%")
print("executed warp-instructions by opcode (top 25):")
te = sum(opc.values())
for o, n in opc.most_common(25):
    print(f"  {o:22s} {n:12d} {100.0 * n / te:5.1f}%")
print(f"top {topn} instructions by samples:")
for ns, src, st, ex in sorted(insts, key=lambda t: -t[0])[:topn]:
    print(f"  {ns:6d}  {src[:70]:70s} {dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])}")
