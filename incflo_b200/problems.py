"""Synthetic inputs of the benchmark configurations.

Closed-form fields taken from the reference's initial conditions:
  taylor_green   src/prob/prob_init_fluid.cpp:243-263   (probtype 1)
  rayleigh_taylor src/prob/prob_init_fluid.cpp:609-665  (probtype 5)
  double_shear_layer src/prob/prob_init_fluid.cpp:869-895 (probtype 21)
plus a non-solenoidal perturbation grad(psi), psi = sin(2 pi x) sin(4 pi y) cos(2 pi z)/(4 pi),
so that the projection has work to do.  Domain is [0,1]^3 scaled so that dx = dy = dz.
All generators return torch float64 tensors on `device`, shaped (ncomp, nz+2ng, ny+2ng, nx+2ng).
"""
import math

import torch


_SLAB = None  # (zlo, nzl): generate only the cell planes [zlo, zlo+nzl) of the global domain n


class slab:
    """with problems.slab(zlo, nzl): ...  -> the generators return that z slab of the global fields"""

    def __init__(self, zlo, nzl):
        self.s = (int(zlo), int(nzl))

    def __enter__(self):
        global _SLAB
        self.prev, _SLAB = _SLAB, self.s

    def __exit__(self, *a):
        global _SLAB
        _SLAB = self.prev


def _local(n):
    """shape of the generated block: the global n with the slab's plane count in z"""
    return (n[0], n[1], _SLAB[1] if _SLAB else n[2])


def _centres(n, device):
    nx, ny, nz = n
    h = 1.0 / nx
    zlo, nzl = _SLAB if _SLAB else (0, nz)
    x = (torch.arange(nx, dtype=torch.float64, device=device) + 0.5) * h
    y = (torch.arange(ny, dtype=torch.float64, device=device) + 0.5) * h
    z = (torch.arange(zlo, zlo + nzl, dtype=torch.float64, device=device) + 0.5) * h
    return x[None, None, :], y[None, :, None], z[:, None, None], h


def grad_psi(n, device, neumann_z=False):
    """grad of psi = sin(2 pi x) sin(4 pi y) cos(2 pi z) / (4 pi); with Neumann walls in z the
    z-dependence cos(2 pi z/Lz) already has zero normal derivative at z = 0, Lz."""
    x, y, z, h = _centres(n, device)
    tp = 2.0 * math.pi
    lz = n[2] * h
    kz = tp / lz
    a = 1.0 / (4.0 * math.pi)
    gx = a * tp * torch.cos(tp * x) * torch.sin(2 * tp * y) * torch.cos(kz * z)
    gy = a * 2 * tp * torch.sin(tp * x) * torch.cos(2 * tp * y) * torch.cos(kz * z)
    gz = -a * kz * torch.sin(tp * x) * torch.sin(2 * tp * y) * torch.sin(kz * z)
    return torch.stack([gx.expand(*_local(n)[::-1]), gy.expand(*_local(n)[::-1]), gz.expand(*_local(n)[::-1])])


def _with_ghosts(v, ng):
    if ng == 0:
        return v.contiguous()
    c, nz, ny, nx = v.shape
    out = torch.zeros((c, nz + 2 * ng, ny + 2 * ng, nx + 2 * ng), dtype=v.dtype, device=v.device)
    out[:, ng:ng + nz, ng:ng + ny, ng:ng + nx] = v
    return out


def taylor_green(n, ng=1, device="cpu", perturb=True):
    x, y, z, h = _centres(n, device)
    tp = 2.0 * math.pi
    u = (torch.sin(tp * x) * torch.cos(tp * y)).expand(*_local(n)[::-1])
    v = (-torch.cos(tp * x) * torch.sin(tp * y)).expand(*_local(n)[::-1])
    w = torch.zeros(_local(n)[::-1], dtype=torch.float64, device=device)
    vel = torch.stack([u, v, w])
    if perturb:
        vel = vel + grad_psi(n, device)
    return _with_ghosts(vel, ng)


def double_shear_layer(n, ng=1, device="cpu", perturb=True):
    x, y, z, h = _centres(n, device)
    tp = 2.0 * math.pi
    u = torch.tanh(30.0 * (0.25 - torch.abs(y - 0.5))).expand(*_local(n)[::-1])
    v = (0.05 * torch.sin(tp * x)).expand(*_local(n)[::-1])
    w = torch.zeros(_local(n)[::-1], dtype=torch.float64, device=device)
    vel = torch.stack([u, v, w])
    if perturb:
        vel = vel + grad_psi(n, device)
    return _with_ghosts(vel, ng)


def rayleigh_taylor_density(n, ngd=0, device="cpu"):
    """rho in [0.5, 2], tanh interface of width 0.005 at z ~ 0.5 Lz (sigma contrast 4:1)."""
    x, y, z, h = _centres(n, device)
    lx, ly, lz = n[0] * h, n[1] * h, n[2] * h
    rho_1, rho_2, width = 0.5, 2.0, 0.005
    r2d = torch.clamp(torch.hypot(x - 0.5 * lx, y - 0.5 * ly), max=0.5 * lx)
    pert = 0.5 * lz - 0.01 * torch.cos(2.0 * math.pi * r2d / lx)
    rho = rho_1 + 0.5 * (rho_2 - rho_1) * (1.0 + torch.tanh((z - pert) / width))
    rho = rho.expand(*_local(n)[::-1]).contiguous()
    if ngd == 0:
        return rho
    nl = _local(n)
    out = torch.ones((nl[2] + 2 * ngd, nl[1] + 2 * ngd, nl[0] + 2 * ngd), dtype=torch.float64, device=device)
    out[ngd:ngd + nl[2], ngd:ngd + nl[1], ngd:ngd + nl[0]] = rho
    return out


def rayleigh_taylor_velocity(n, ng=1, device="cpu", case="b"):
    """case 'a': the InitialPressureProjection input u = g = (0,0,-0.1) in valid cells and one
    ghost layer (src/setup/init.cpp:533-560); case 'b': Taylor-Green + grad(psi) (walls in z)."""
    if case == "a":
        assert _SLAB is None
        vel = torch.zeros((3, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng), dtype=torch.float64, device=device)
        lo = ng - 1
        vel[2, lo:ng + n[2] + 1, lo:ng + n[1] + 1, lo:ng + n[0] + 1] = -0.1
        return vel
    x, y, z, h = _centres(n, device)
    tp = 2.0 * math.pi
    u = (torch.sin(tp * x) * torch.cos(tp * y)).expand(*_local(n)[::-1])
    v = (-torch.cos(tp * x) * torch.sin(tp * y)).expand(*_local(n)[::-1])
    w = torch.zeros(_local(n)[::-1], dtype=torch.float64, device=device)
    vel = torch.stack([u, v, w]) + grad_psi(n, device, neumann_z=True)
    return _with_ghosts(vel, ng)


# BC codes: 0 periodic, 1 Neumann, 2 Dirichlet, 3 inflow
CONFIGS = {
    # BASELINE.json configs[0]: taylor_green_vortices, periodic, constant density
    "tgv": dict(bclo=(0, 0, 0), bchi=(0, 0, 0), var=False),
    # configs[1]: rayleigh_taylor, variable density, periodic x,y + slip walls in z
    "rt": dict(bclo=(0, 0, 1), bchi=(0, 0, 1), var=True),
    # configs[2]: double_shear_layer_x, periodic, constant density
    "dsl": dict(bclo=(0, 0, 0), bchi=(0, 0, 0), var=False),
}


def make(config, N, ng=1, device="cpu"):
    """returns dict(vel, sigma|None, const_sigma, bclo, bchi, n, dx) for projector-level tests."""
    n = (N, N, N)
    cfg = CONFIGS[config]
    dt = 0.45 / N  # cfl 0.45, |u| ~ 1  (test_no_eb_3d/benchmark.taylor_green_vortices:12)
    if config == "tgv":
        vel = taylor_green(n, ng, device)
        sigma, cs = None, dt / 1.0
    elif config == "dsl":
        vel = double_shear_layer(n, ng, device)
        sigma, cs = None, dt / 1.0
    else:
        vel = rayleigh_taylor_velocity(n, ng, device, "b")
        sigma, cs = (dt / rayleigh_taylor_density(n, 0, device)).contiguous(), 1.0
    return dict(vel=vel, sigma=sigma, const_sigma=cs, bclo=cfg["bclo"], bchi=cfg["bchi"], n=n, dx=(1.0 / N,) * 3, dt=dt)
