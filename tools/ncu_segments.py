"""Where does a kernel spend its INSTRUCTIONS?  Splits the SASS of one captured kernel (ncu --set full --import-source on) into equal
segments and prints, per segment, the share of executed warp instructions, the share of stall samples and the most executed opcodes --
the view that showed k_interp_tile to be issue bound in its staging index arithmetic (profiles/r2e_ncu_full_interp_256_rt.txt).
Usage: python tools/ncu_segments.py report.ncu-rep [segments]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nseg = int(sys.argv[2]) if len(sys.argv) > 2 else 24
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr_at = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
print(rows[hdr_at - 1][1] if hdr_at > 0 and len(rows[hdr_at - 1]) > 1 else "")
hdr, data = rows[hdr_at], [r for r in rows[hdr_at + 1:] if len(r) > 5]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data) or 1
ts = sum(int(r[isamp]) for r in data) or 1
print(f"{len(data)} SASS instructions, {tot} warp instructions executed, {ts} stall samples")
seg = max(1, (len(data) + nseg - 1) // nseg)
for s in range(0, len(data), seg):
    chunk = data[s:s + seg]
    e = sum(int(r[ia]) for r in chunk)
    sm = sum(int(r[isamp]) for r in chunk)
    ops = {}
    for r in chunk:
        w = r[isrc].split()
        op = (w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "?"))
        ops[op] = ops.get(op, 0) + int(r[ia])
    top = sorted(ops.items(), key=lambda x: -x[1])[:5]
    print(f"  SASS {s:5d}-{min(s + seg, len(data)):5d}  executed {100 * e / tot:5.1f} %  samples {100 * sm / ts:5.1f} %   " +
          " ".join(f"{k}:{100 * v / tot:.1f}" for k, v in top))
