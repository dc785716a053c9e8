"""MAC projection timing on one GPU: rayleigh_taylor-like variable density (4:1), periodic x/y + walls z, N^3 cells,
face velocity = smooth random field, device-resident arrays.  Usage: python tools/mac_bench.py [N] [steps]
Prints one JSON line: ms per projection, V-cycles, and the whole-solve algorithmic bandwidth (bytes of DESIGN.md section 11
/ time / measured HBM peak)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incflo_b200 import mac_projector as mp

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
peak = 6545.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
n = (N, N, N)
z = (torch.arange(N, device=dev, dtype=torch.float64) + 0.5) / N
rho = 1.0 + 3.0 * 0.5 * (1.0 + torch.tanh((z - 0.5) / 0.05))
rho = rho[:, None, None].expand(N, N, N).contiguous()
dt = 0.01
bx = (dt / (0.5 * (rho + torch.roll(rho, 1, 2))))
bx = torch.cat([bx, bx[:, :, :1]], 2).contiguous()
by = (dt / (0.5 * (rho + torch.roll(rho, 1, 1))))
by = torch.cat([by, by[:, :1]], 1).contiguous()
rz = torch.cat([rho[:1], rho, rho[-1:]], 0)
bz = (dt / (0.5 * (rz[:-1] + rz[1:]))).contiguous()


def smooth(a):
    for ax in range(3):
        a = 0.5 * a + 0.25 * (torch.roll(a, 1, ax) + torch.roll(a, -1, ax))
    return a


u0 = smooth(torch.randn((N, N, N + 1), device=dev, dtype=torch.float64, generator=g)); u0[:, :, -1] = u0[:, :, 0]
v0 = smooth(torch.randn((N, N + 1, N), device=dev, dtype=torch.float64, generator=g)); v0[:, -1] = v0[:, 0]
w0 = smooth(torch.randn((N + 1, N, N), device=dev, dtype=torch.float64, generator=g)); w0[0] = 0; w0[-1] = 0
proj = mp.MacProjector(n, (1.0 / N,) * 3, (0, 0, 1), (0, 0, 1))
proj.updateCoeffs([bx, by, bz])
phi = torch.zeros((N, N, N), device=dev, dtype=torch.float64)
times = []
for s in range(K + 2):
    u, v, w = u0.clone(), v0.clone(), w0.clone()
    torch.cuda.synchronize()
    st = proj.project(u, v, w, 1e-11, 1e-14, mac_phi=phi)
    if s >= 2:
        times.append(st.ms_total)
ms = sum(times) / len(times)
# algorithmic bytes per cell and V-cycle: 8 half-sweeps x 40 + residual 48 + restriction 9 + interpolation 17 on every level, + the top-level residual
# norm (40: nothing stored, no sol += cor in the direct form; B200MAC_TOP_DIRECT=0 adds 24 + 8)
direct = os.environ.get("B200MAC_TOP_DIRECT", "1") != "0"
per_cell = (8 * 40 + 48 + 9 + 17) * 8.0 / 7.0 + (40 if direct else 24 + 48)
total = per_cell * N ** 3 * st.iters + (24 + 8 + 48 + 8 + 3 * 24) * N ** 3
print(json.dumps({"metric": "mac_projection_Mcell_updates_per_s", "value": N ** 3 / ms / 1e3, "unit": "Mcell-updates/s", "n": N,
                  "ms_per_projection": ms, "ms_solve": st.ms_solve, "vcycles": st.iters, "resid_over_bnorm": st.resnorm / max(st.rhsnorm, st.resnorm0),
                  "launches": st.launches, "whole_solve": {"algorithmic_bytes": total, "achieved_GBs": total / ms / 1e6, "frac_of_measured_peak": total / ms / 1e6 / peak}}))
