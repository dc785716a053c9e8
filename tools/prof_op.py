"""Launch one V-cycle building block on level `lev` a few times (for ncu captures).
Usage: python tools/prof_op.py N cfg op [lev] [reps]   op in smooth|residual|restrict|interp|vcycle"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from incflo_b200 import nodal_projector as npj, problems

N = int(sys.argv[1]); cfgname = sys.argv[2]; opname = sys.argv[3]
lev = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
cfg = problems.make(cfgname, N, ng=1, device="cuda:0")
proj = npj.NodalProjector(cfg["vel"], cfg["sigma"], cfg["const_sigma"],
                          dict(n_cell=cfg["n"], dx=cfg["dx"], is_periodic=[b == 0 for b in cfg["bclo"]]), ng=1)
proj.setDomainBC(cfg["bclo"], cfg["bchi"])
proj.set_sigma(cfg["sigma"], cfg["const_sigma"])
op = dict(smooth=npj.OP_SMOOTH, residual=npj.OP_RESIDUAL, restrict=npj.OP_RESTRICT, interp=npj.OP_INTERP,
          vcycle=npj.OP_VCYCLE)[opname]
ms = proj.time_op(lev, op, 1, reps=reps)
print(f"{opname} lev{lev} N={N} {cfgname}: {ms * 1e3:.1f} us per launch")
