"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals and shares."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    rd = csv.reader(l for l in f if l.startswith('"'))
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    ig = hdr.index("Grid Size") if "Grid Size" in hdr else None
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        v_us = v / 1e3 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1e3)
        rows.append((r[ik].split("(")[0][:50], r[ig] if ig is not None else "", v_us))
tot = sum(r[2] for r in rows)
agg = defaultdict(lambda: [0, 0.0])
for k, g, v in rows:
    agg[k][0] += 1; agg[k][1] += v
print(f"{len(rows)} launches, total {tot / 1e3:.3f} ms (cold-cache serialised ncu times: compare SHARES)")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:52s} n={n:5d}  {v / 1e3:9.3f} ms  {100 * v / tot:5.1f} %  avg {v / n:9.1f} us")
if len(sys.argv) > 2:  # big-grid launches only
    big = defaultdict(lambda: [0, 0.0])
    for k, g, v in rows:
        big[(k, g)][0] += 1; big[(k, g)][1] += v
    for (k, g), (n, v) in sorted(big.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"  {k:44s} {g:18s} n={n:5d} {v / 1e3:9.3f} ms {100 * v / tot:5.1f} % avg {v / n:9.1f} us")
