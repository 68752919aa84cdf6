#!/usr/bin/env python
"""Per-kernel launch counts, total time and share from an `ncu --metrics gpu__time_duration.sum --csv`
launch list (cold-cache, serialised launches: compare SHARES with bench.py's breakdown, not absolutes)."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    tot[name][0] += 1
    tot[name][1] += float(r[-1]) * 1e-6
allms = sum(v[1] for v in tot.values())
print("kernel, launches, total_ms, share (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache "
      "serialised launches: compare shares)")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k}, {n}, {ms:.3f}, {ms / allms:.4f}")
