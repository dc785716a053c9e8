"""BASELINE configs[3] at full size: bouss_bubble-like composite projection, 256^3 base level (periodic x/y,
walls z, constant density) + the central 128^3 coarse cells refined by 2 (a 256^3 fine level).
u = (0, 0, -0.5 T) (src/prob/prob_init_fluid.cpp:797-808, Boussinesq buoyancy kick) + grad(psi).
Usage: python tools/composite_bench.py [N] [reps]"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incflo_b200 import nodal_projector as npj

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda:0"
ng = 3
clo = (N // 4,) * 3
chi = (3 * N // 4 - 1,) * 3
nf = N


def field(n, h, off):
    ar = lambda m, o: (torch.arange(m, dtype=torch.float64, device=dev) + 0.5) * h + o
    x = ar(n, off)[None, None, :]; y = ar(n, off)[None, :, None]; z = ar(n, off)[:, None, None]
    tp = 2 * math.pi
    r2 = (x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.35) ** 2
    T = 0.5 * torch.exp(-r2 / 0.01)                       # warm bubble
    gx = -tp * torch.sin(tp * x) * torch.cos(tp * y) * torch.cos(math.pi * z)
    gy = -tp * torch.cos(tp * x) * torch.sin(tp * y) * torch.cos(math.pi * z)
    gz = -math.pi * torch.cos(tp * x) * torch.cos(tp * y) * torch.sin(math.pi * z) - 0.5 * T
    v = torch.zeros((3, n + 2 * ng, n + 2 * ng, n + 2 * ng), dtype=torch.float64, device=dev)
    v[0, ng:-ng, ng:-ng, ng:-ng] = 0.05 * gx.expand(n, n, n)
    v[1, ng:-ng, ng:-ng, ng:-ng] = 0.05 * gy.expand(n, n, n)
    v[2, ng:-ng, ng:-ng, ng:-ng] = 0.05 * gz.expand(n, n, n)
    return v


vel_in = [field(N, 1.0 / N, 0.0), field(nf, 0.5 / N, clo[0] / N)]
gp = [torch.zeros((3, N, N, N), dtype=torch.float64, device=dev), torch.zeros((3, nf, nf, nf), dtype=torch.float64, device=dev)]
p = [torch.zeros((N + 1,) * 3, dtype=torch.float64, device=dev), torch.zeros((nf + 1,) * 3, dtype=torch.float64, device=dev)]
cp = npj.CompositeProjection((N, N, N), (1.0 / N,) * 3, (0, 0, 1), (0, 0, 1), clo, chi)
dt = 0.45 / N
for it in range(reps + 1):
    vel = [v.clone() for v in vel_in]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = cp.apply_nodal_projection(vel, (ng, ng), gp, p, density=None, ro_0=1.0, scaling_factor=dt)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    cells = N ** 3 - (N // 2) ** 3 + nf ** 3       # composite cells: uncovered coarse + fine
    print(f"composite {N}^3 base + {nf}^3 fine: iters={st.iters} status={st.status} resid/bnorm={st.resnorm / max(st.rhsnorm, st.resnorm0):.3e} "
          f"ms_total={st.ms_total:.2f} ms_solve={st.ms_solve:.2f} wall={wall * 1e3:.2f} ms launches={st.launches} "
          f"Mcell/s={cells / st.ms_total / 1e3:.1f}")
cp.close()
