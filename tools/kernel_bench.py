"""Per-kernel timing of the V-cycle building blocks on level 0 (CUDA events, back-to-back
launches on arrays larger than L2 at N >= 256).  Usage: python tools/kernel_bench.py N [var]"""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incflo_b200 import nodal_projector as npj, problems

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfgname = sys.argv[2] if len(sys.argv) > 2 else "rt"
peak = 6539.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
cfg = problems.make(cfgname, N, ng=1, device="cuda:0")
var = cfg["sigma"] is not None
proj = npj.NodalProjector(cfg["vel"], cfg["sigma"], cfg["const_sigma"],
                          dict(n_cell=cfg["n"], dx=cfg["dx"], is_periodic=[b == 0 for b in cfg["bclo"]]), ng=1)
proj.setDomainBC(cfg["bclo"], cfg["bchi"])
proj.set_sigma(cfg["sigma"], cfg["const_sigma"])
st = proj.project(1e-11, 1e-14)
print(f"N={N} cfg={cfgname} var={var} iters={st.iters} ms_total={st.ms_total:.2f} ms_solve={st.ms_solve:.2f} "
      f"launches={st.launches} Mcells/s={N**3 / st.ms_total / 1e3:.1f}")
for lev in range(proj.nlevels() if os.environ.get('KB_ALL') else min(3, proj.nlevels())):
    n, nn = proj.level_dims(lev)
    nodes = nn[0] * nn[1] * nn[2]
    bytes_per = {"smooth": 32 if var else 24, "residual": 32 if var else 24, "restrict": 9, "interp": 25 if var else 17}
    for name, op, arg in (("smooth", npj.OP_SMOOTH, 2), ("residual", npj.OP_RESIDUAL, 0), ("restrict", npj.OP_RESTRICT, 0),
                          ("interp", npj.OP_INTERP, 0)):
        if lev + 1 >= proj.nlevels() and name in ("restrict", "interp"):
            continue
        ms = proj.time_op(lev, op, arg, reps=10)
        if name == "smooth":
            ms /= 2
        gbs = nodes * bytes_per[name] / ms / 1e6
        print(f"  lev{lev} {nn} {name:9s} {ms * 1e3:9.1f} us  {gbs:8.1f} GB/s algorithmic  {gbs / peak:6.1%} of measured HBM peak")
ms = proj.time_op(0, npj.OP_VCYCLE, 0, reps=3)
print(f"  vcycle (no graph) {ms:.3f} ms")
