"""EB nodal projection timing on one GPU: BASELINE configs[4], test_3d/benchmark.channel_cylinder-x scaled to NX x NY x NZ cells
(default 512 x 128 x 128, dx = 1/320: domain 1.6 x 0.4 x 0.4; cylinder r = 0.05000001 along z at (0.151, 0.2); mass inflow x-lo
with the parabolic profile of probtype 31, pressure outflow x-hi, no-slip walls y, periodic z; constant density ro_0 = 1).
One step = the initial nodal projection of u = (1, 0, 0) (scaling_factor 1) to rtol 1e-11, device-resident arrays.
Usage: python tools/eb_bench.py [NX NY NZ] [steps].  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from incflo_b200 import eb_geometry as eg, eb_projector as ebp

n = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (512, 128, 128)
K = int(sys.argv[4]) if len(sys.argv) > 4 else 5
VAR = len(sys.argv) > 5 and sys.argv[5] == "var"      # variable density: sigma = dt / rho as a cell array (4:1), e.g. test_3d/benchmark.eb_flow_density
h = 0.4 / n[1]
peak = 6545.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
t0 = time.time()
geom = eg.cylinder(n, h, 0.05000001, (0.151, 0.2, 0.0), direction=2, small_vfrac=1e-6)
t_geom = time.time() - t0
dev = torch.device("cuda:0")
vel0 = np.zeros((3, n[2] + 2, n[1] + 2, n[0] + 2))
vel0[0, 1:-1, 1:-1, 1:-1] = (geom.vfrac > 0)                       # ic_u = 1 in the fluid
y = (np.arange(n[1]) + 0.5) / n[1]
vel0[0, 1:-1, 1:-1, 0] = (6.0 * y * (1.0 - y))[None, :]           # IncfloVelFill, probtype 31 (src/prob/prob_bc.H:62-66)
bclo, bchi = (3, 1, 0), (2, 1, 0)
proj = ebp.EBNodalProjector(n, (h,) * 3, bclo, bchi, geom.vfrac, geom.intg)
tv0 = torch.from_numpy(vel0).to(dev)
phi = torch.zeros((n[2] + 1, n[1] + 1, n[0] + 1), device=dev, dtype=torch.float64)
gphi = torch.zeros((3, n[2], n[1], n[0]), device=dev, dtype=torch.float64)
sigma = 1.0
if VAR:
    zz = (torch.arange(n[1], device=dev, dtype=torch.float64) + 0.5) / n[1]
    sigma = (1.0 / (1.0 + 1.5 * (1.0 + torch.tanh((zz - 0.5) / 0.1))))[None, :, None].expand(n[2], n[1], n[0]).contiguous()
times, solve = [], []
for s in range(K + 2):
    tv = tv0.clone()
    torch.cuda.synchronize()
    st = proj.project(tv, sigma, 1e-11, 1e-14, phi=phi, gphi=gphi)
    if s >= 2:
        times.append(st.ms_total); solve.append(st.ms_solve)
ms, mss = sum(times) / len(times), sum(solve) / len(solve)
# per-level smoother / residual times (one smooth call = smooth_num_sweeps sweeps of 8 colours)
levels = []
for lev in range(proj.nlevels()):
    nc_, nn_ = proj.level_dims(lev)
    nnod = nn_[0] * nn_[1] * nn_[2]
    t_s = proj.time_op(lev, 0, 1, 10) / 4.0
    t_r = proj.time_op(lev, 1, 1, 10)
    levels.append({"lev": lev, "nodes": nnod, "us_per_sweep": 1e3 * t_s, "us_per_residual": 1e3 * t_r,
                   "sweep_GBs_at_240B_per_node": 240.0 * nnod / t_s / 1e6, "sweep_GBs_at_49B_per_node": 49.0 * nnod / t_s / 1e6})
ncell = n[0] * n[1] * n[2]
nnode = (n[0] + 1) * (n[1] + 1) * n[2]
# algorithmic bytes per node and V-cycle: 16 sweeps x (27 coefficients + rhs + phi in/out = 240 B) + residual 240 + restriction 9 + interpolation 17,
# x 8/7 for the hierarchy, + top level axpy 24 + residual 240; set-up: stencil build 19 x 8 in + 27 x 8 out, Galerkin products 27 x 8 in / 8 + out
per_node = (16 * 240 + 240 + 9 + 17) * 8.0 / 7.0 + 24 + 240
total = per_node * nnode * st.iters + (19 * 8 + 27 * 8 + 27 * 8 * 2.0 / 7.0) * nnode
u = tv[:, 1:-1, 1:-1, 1:-1]
print(json.dumps({"metric": "eb_nodal_projection_Mcell_updates_per_s", "value": ncell / ms / 1e3, "unit": "Mcell-updates/s", "n": n,
                  "workload": "channel_cylinder-x (BASELINE configs[4]) EB cylinder, inflow / outflow / walls / periodic z, " + ("variable density 4:1 across y" if VAR else "constant density"),
                  "ms_per_projection": ms, "ms_solve": mss, "ms_setup_and_update": ms - mss, "vcycles": st.iters, "nlevels": st.nlevels,
                  "bottom_iters": st.bottom_iters, "resid_over_bnorm": st.resnorm / max(st.rhsnorm, st.resnorm0), "launches": st.launches,
                  "cut_cells": int(geom.cut_mask().sum()), "covered_cells": int((geom.vfrac == 0).sum()), "geometry_s": t_geom,
                  "max_u_after": float(u.abs().max()), "levels": levels,
                  "whole_solve": {"algorithmic_bytes": total, "achieved_GBs": total / ms / 1e6, "frac_of_measured_peak": total / ms / 1e6 / peak}}))
