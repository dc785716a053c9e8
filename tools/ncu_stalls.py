"""Per-kernel stall-reason totals and the hottest SASS instructions from an .ncu-rep source page
(read here, no GPU):  python tools/ncu_stalls.py file.ncu-rep [kernel-substring] [top-n]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kernels, cur = [], None
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = dict(name=row[1], hdr=None, rows=[])
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None:
        cur["rows"].append(row)
seen = set()
for k in kernels:
    if want not in k["name"] or k["name"] in seen:
        continue
    seen.add(k["name"])
    h = {n: i for i, n in enumerate(k["hdr"])}
    stall_cols = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = defaultdict(float)
    samples = 0.0
    for r in k["rows"]:
        samples += float(r[h["# Samples"]] or 0)
        for c in stall_cols:
            tot[c] += float(r[h[c]] or 0)
    print("===", k["name"][:90], f"({int(samples)} samples, {len(k['rows'])} SASS instructions)")
    for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]:
        print(f"   {c:24s} {100 * v / max(samples, 1):5.1f} %")
    print("   hottest instructions (share of samples, dominant stall):")
    for r in sorted(k["rows"], key=lambda r: -float(r[h["# Samples"]] or 0))[:topn]:
        s = float(r[h["# Samples"]] or 0)
        dom = max(stall_cols, key=lambda c: float(r[h[c]] or 0))
        print(f"   {100 * s / samples:5.2f} %  {dom[6:]:14s} {r[h['Source']].strip()[:90]}")
