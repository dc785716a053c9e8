"""MAC projection (b200mac_*): the CUDA kernels through the C ABI against the numpy oracle (oracle/mac_oracle.py), building
block by building block and end to end, and against the independent SciPy fixtures tests/golden/mac_*.npz.
Tolerances: building blocks 1e-12 relative (same arithmetic, different summation order); projections 1e-9 relative L2 at
rtol 1e-11 / 1e-12 (north_star's bar for the nodal projection, applied to this operator too)."""
import glob
import os

import numpy as np
import pytest

from oracle import mac_oracle as mo

pytestmark = pytest.mark.gpu

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mac_*.npz")))
CASES = [
    ("periodic", (32, 32, 32), (1 / 32,) * 3, (0, 0, 0), (0, 0, 0)),
    ("rt_walls_z", (32, 16, 24), (1 / 32,) * 3, (0, 0, 1), (0, 0, 1)),
    ("channel_inflow_outflow", (48, 16, 16), (1 / 48,) * 3, (3, 1, 0), (2, 1, 0)),
    ("aniso_dirichlet", (16, 24, 20), (0.1, 0.07, 0.05), (2, 1, 1), (2, 1, 2)),
    ("odd_periodic_bottom", (24, 12, 12), (1 / 24,) * 3, (0, 0, 1), (0, 0, 1)),     # coarsens to 6 x 3 x 3: odd periodic extent
]


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def _bc(b):
    return tuple(1 if x == 3 else x for x in b)      # inflow -> Neumann (get_mac_projection_bc)


def _beta(n, rng, var):
    if not var:
        return 0.37
    return [rng.uniform(0.5, 2.0, size=(n[2], n[1], n[0] + 1)), rng.uniform(0.5, 2.0, size=(n[2], n[1] + 1, n[0])),
            rng.uniform(0.5, 2.0, size=(n[2] + 1, n[1], n[0]))]


def _make(case, var, seed=0):
    from incflo_b200 import mac_projector as mp
    name, n, dx, bclo, bchi = case
    rng = np.random.default_rng(seed)
    beta = _beta(n, rng, var)
    if var:   # periodic faces: one coefficient per physical face
        for d, ax in ((0, 2), (1, 1), (2, 0)):
            if bclo[d] == 0:
                sl = lambda s: tuple(s if a == ax else slice(None) for a in range(3))
                beta[d][sl(n[d])] = beta[d][sl(0)]
    mg = mo.MG(mo.Params(n, dx, _bc(bclo), _bc(bchi)), beta)
    proj = mp.MacProjector(n, dx, bclo, bchi)
    proj.updateCoeffs(beta if var else 0.37)
    return mg, proj, rng


@pytest.mark.parametrize("var", [False, True])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_mac_building_blocks(case, var):
    from incflo_b200 import mac_projector as mp
    mg, proj, rng = _make(case, var)
    assert proj.nlevels() == len(mg.lv)
    for lev in range(len(mg.lv)):
        n = mg.lv[lev].n
        assert proj.level_dims(lev) == tuple(n)
        shp = (n[2], n[1], n[0])
        x, r = rng.standard_normal(shp), rng.standard_normal(shp)
        assert rel(proj.level_op(lev, mp.OP_RESIDUAL, 0, x, r), mg.residual(lev, x, r)) < 1e-12
        got = proj.level_op(lev, mp.OP_SMOOTH, 2, x, r)
        assert rel(got, mg.smooth(lev, x.copy(), r, 2)) < 1e-12
        if lev + 1 < len(mg.lv):
            nc = mg.lv[lev + 1].n
            assert rel(proj.level_op(lev, mp.OP_RESTRICT, 0, x, out_lev=lev + 1), mg.restrict(x)) < 1e-13
            c = rng.standard_normal((nc[2], nc[1], nc[0]))
            assert rel(proj.level_op(lev, mp.OP_INTERP, 0, x, c), mg.interp_add(x.copy(), c)) < 1e-14
    nb = mg.lv[-1].n
    b = rng.standard_normal((nb[2], nb[1], nb[0]))
    got = proj.level_op(len(mg.lv) - 1, mp.OP_BOTTOM, 0, None, b)
    want = mg.bottom_solve(np.zeros_like(b), b.copy())
    assert rel(got, want) < 1e-8     # BiCGStab: reductions in a different order
    proj.close()


@pytest.mark.parametrize("host", [True, False], ids=["host_ptrs", "device_ptrs"])
@pytest.mark.parametrize("var", [False, True])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_mac_project_parity(case, var, host):
    import torch
    name, n, dx, bclo, bchi = case
    mg, proj, rng = _make(case, var, seed=3)
    u, v, w = rng.standard_normal((n[2], n[1], n[0] + 1)), rng.standard_normal((n[2], n[1] + 1, n[0])), rng.standard_normal((n[2] + 1, n[1], n[0]))
    vel = [u, v, w]
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        for arr in vel:
            pass
        sl = lambda s: tuple(s if a == ax else slice(None) for a in range(3))
        if bclo[d] == 0:
            vel[d][sl(n[d])] = vel[d][sl(0)]
    beta = mg.lv[0].b if var else 0.37
    ou, ov, ow = u.copy(), v.copy(), w.copy()
    r = mo.project(mg.p, ou, ov, ow, beta, 1e-11, 1e-14)
    assert r["stats"]["status"] == 0
    conv = (lambda a: a.copy()) if host else (lambda a: torch.from_numpy(a.copy()).cuda())
    gu, gv, gw = conv(u), conv(v), conv(w)
    gphi = conv(np.zeros((n[2], n[1], n[0])))
    st = proj.project(gu, gv, gw, 1e-11, 1e-14, mac_phi=gphi)
    back = (lambda a: a) if host else (lambda a: a.cpu().numpy())
    assert st.status == 0 and abs(st.iters - r["stats"]["iters"]) <= 1
    assert st.resnorm <= max(1e-14, 1e-11 * max(st.rhsnorm, st.resnorm0))
    assert abs(st.rhsnorm - r["stats"]["rhsnorm"]) <= 1e-12 * r["stats"]["rhsnorm"]
    singular = all(b != 2 for b in bclo + bchi)
    a, b = back(gphi), r["phi"]
    if singular:
        a, b = a - a.mean(), b - b.mean()
    assert rel(a, b) < 1e-9
    assert rel(back(gu), ou) < 1e-9 and rel(back(gv), ov) < 1e-9 and rel(back(gw), ow) < 1e-9
    # the projected face velocity is discretely divergence-free (up to the solver tolerance)
    uu, vv, ww = back(gu), back(gv), back(gw)
    div = (uu[:, :, 1:] - uu[:, :, :-1]) / dx[0] + (vv[:, 1:] - vv[:, :-1]) / dx[1] + (ww[1:] - ww[:-1]) / dx[2]
    if singular:
        div = div - div.mean()
    assert np.abs(div).max() <= 2e-11 * max(st.rhsnorm, st.resnorm0)
    proj.close()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_mac_cuda_reproduces_golden(path):
    from incflo_b200 import mac_projector as mp
    z = np.load(path)
    n = tuple(int(x) for x in z["n"]); dx = tuple(float(x) for x in z["dx"])
    bclo = tuple(int(x) for x in z["bclo"]); bchi = tuple(int(x) for x in z["bchi"])
    proj = mp.MacProjector(n, dx, bclo, bchi)
    proj.updateCoeffs([z["bx"].copy(), z["by"].copy(), z["bz"].copy()])
    u, v, w = z["u_in"].copy(), z["v_in"].copy(), z["w_in"].copy()
    phi = np.zeros((n[2], n[1], n[0]))
    st = proj.project(u, v, w, 1e-12, 0.0, mac_phi=phi)
    assert st.status == 0
    singular = all(b != 2 for b in bclo + bchi)
    a, b = (phi - phi.mean(), z["phi"] - z["phi"].mean()) if singular else (phi, z["phi"])
    assert rel(a, b) < 1e-9
    assert rel(u, z["u_out"]) < 1e-9 and rel(v, z["v_out"]) < 1e-9 and rel(w, z["w_out"]) < 1e-9
    proj.close()


def test_mac_initial_guess_and_update_beta():
    """project(mac_phi, ...) starts from mac_phi (:287-292); updateBeta(const) rebuilds the coefficients on every level"""
    from incflo_b200 import mac_projector as mp
    n, dx = (32, 32, 32), (1 / 32,) * 3
    bclo = bchi = (0, 0, 1)
    rng = np.random.default_rng(5)
    u, v, w = rng.standard_normal((32, 32, 33)), rng.standard_normal((32, 33, 32)), rng.standard_normal((33, 32, 32))
    u[:, :, -1] = u[:, :, 0]; v[:, -1] = v[:, 0]; w[0] = 0; w[-1] = 0
    proj = mp.MacProjector(n, dx, bclo, bchi)
    proj.updateCoeffs(0.5)
    phi = np.zeros((32, 32, 32))
    a = [x.copy() for x in (u, v, w)]
    it1 = proj.project(*a, 1e-11, 1e-14, mac_phi=phi).iters
    # the converged phi as the initial guess: nothing left to do
    b = [x.copy() for x in (u, v, w)]
    it2 = proj.project(*b, 1e-10, 1e-14, mac_phi=phi, use_phi_as_guess=True).iters
    assert it1 > 3 and it2 == 0
    assert rel(b[0], a[0]) < 1e-9
    # beta -> 2 beta: phi halves, the projected velocity stays
    proj.updateCoeffs(1.0)
    phi2 = np.zeros_like(phi)
    c = [x.copy() for x in (u, v, w)]
    proj.project(*c, 1e-11, 1e-14, mac_phi=phi2)
    assert rel((phi2 - phi2.mean()) * 2.0, phi - phi.mean()) < 1e-8 and rel(c[2], a[2]) < 1e-8
    proj.close()


@pytest.mark.parametrize("host", [True, False], ids=["host_ptrs", "device_ptrs"])
@pytest.mark.parametrize("case,max_grid,ng", [(CASES[2], 16, 1), (CASES[1], 8, 0), (CASES[3], 12, 2)], ids=["channel_16", "rt_8", "aniso_ragged_12"])
def test_mac_multibox_equals_single_box(case, max_grid, ng, host):
    """b200mac_set_coeffs_mf / b200mac_project_mf: the MultiFabs of a deck with amr.max_grid_size < domain (every reference deck: 16)
    against the single-box call -- bit for bit on the valid faces and cells; ghost faces are neither read nor written"""
    import torch
    from incflo_b200 import mac_projector as mp
    name, n, dx, bclo, bchi = case
    rng = np.random.default_rng(12)
    beta = _beta(n, rng, True)
    vel = [rng.standard_normal((n[2], n[1], n[0] + 1)), rng.standard_normal((n[2], n[1] + 1, n[0])), rng.standard_normal((n[2] + 1, n[1], n[0]))]
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        if bclo[d] == 0:
            sl = lambda s: tuple(s if a == ax else slice(None) for a in range(3))
            beta[d][sl(n[d])] = beta[d][sl(0)]
            vel[d][sl(n[d])] = vel[d][sl(0)]
    # single box
    p1 = mp.MacProjector(n, dx, bclo, bchi)
    p1.updateCoeffs(beta)
    u1, v1, w1 = (a.copy() for a in vel)
    phi1 = np.zeros((n[2], n[1], n[0]))
    st1 = p1.project(u1, v1, w1, 1e-11, 1e-14, mac_phi=phi1)
    p1.close()
    # multi-box
    to = None if host else (lambda a: torch.from_numpy(a).cuda())
    GARBAGE = 7.25e33
    B = [mp.FaceMultiFab.split(beta[d], n, max_grid, ng, d, to=to, fill=GARBAGE) for d in range(3)]
    U = [mp.FaceMultiFab.split(vel[d], n, max_grid, ng, d, to=to, fill=GARBAGE) for d in range(3)]
    PHI = mp.FaceMultiFab.split(np.zeros((n[2], n[1], n[0])), n, max_grid, ng, -1, to=to, fill=GARBAGE)
    assert len(U[0].boxes) > 1
    p2 = mp.MacProjector(n, dx, bclo, bchi)
    p2.updateCoeffs_mf(*B)
    st2 = p2.project_mf(U[0], U[1], U[2], 1e-11, 1e-14, mac_phi=PHI)
    if not host:
        torch.cuda.synchronize()
    assert st2.status == 0 and st2.iters == st1.iters and st2.resnorm == st1.resnorm
    assert (st2.h2d_bytes > 0) == host
    for got, want in zip(U + [PHI], [u1, v1, w1, phi1]):
        assert np.array_equal(got.assemble(n), want)
    if ng > 0:   # ghost faces untouched
        a = U[0].arrays[0]
        a = a.cpu().numpy() if hasattr(a, "cpu") else a
        assert np.all(a[0] == GARBAGE) and np.all(a[:, :, 0] == GARBAGE)
    p2.close()


def test_mac_caller_stream():
    """b200mac_set_stream: the projection on the caller's stream gives the same bits"""
    import ctypes as C
    import torch
    mg, proj, rng = _make(CASES[1], True, seed=4)
    name, n, dx, bclo, bchi = CASES[1]
    u, v, w = rng.standard_normal((n[2], n[1], n[0] + 1)), rng.standard_normal((n[2], n[1] + 1, n[0])), rng.standard_normal((n[2] + 1, n[1], n[0]))
    u[:, :, -1] = u[:, :, 0]; v[:, -1] = v[:, 0]
    a = [torch.from_numpy(x.copy()).cuda() for x in (u, v, w)]
    st = proj.project(a[0], a[1], a[2], 1e-11, 1e-14)
    it = st.iters
    s = torch.cuda.Stream()
    assert proj._L.b200mac_set_stream(proj._h, C.c_void_p(s.cuda_stream)) == 0
    with torch.cuda.stream(s):
        b = [torch.from_numpy(x.copy()).cuda() for x in (u, v, w)]
        st = proj.project(b[0], b[1], b[2], 1e-11, 1e-14)
    s.synchronize()
    assert st.iters == it and all(torch.equal(x, y) for x, y in zip(a, b))
    assert proj._L.b200mac_set_stream(proj._h, None) == 0
    proj.close()


def test_mac_properties_at_bench_size():
    """the bench configuration (256^3, rayleigh_taylor-like 4:1 density on the faces, periodic x / y + walls z) through properties that need no
    oracle: convergence to rtol, the projected face velocity is discretely divergence-free, linearity in u, phi ~ 1 / beta at fixed u"""
    import torch
    from incflo_b200 import mac_projector as mp
    N = 256
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(3)
    z = (torch.arange(N, device=dev, dtype=torch.float64) + 0.5) / N
    rho = (1.0 + 1.5 * (1.0 + torch.tanh((z - 0.5) / 0.05)))[:, None, None].expand(N, N, N).contiguous()
    dt = 0.01
    bx = dt / (0.5 * (rho + torch.roll(rho, 1, 2))); bx = torch.cat([bx, bx[:, :, :1]], 2).contiguous()
    by = dt / (0.5 * (rho + torch.roll(rho, 1, 1))); by = torch.cat([by, by[:, :1]], 1).contiguous()
    rz = torch.cat([rho[:1], rho, rho[-1:]], 0)
    bz = (dt / (0.5 * (rz[:-1] + rz[1:]))).contiguous()

    def smooth(a):
        for ax in range(3):
            a = 0.5 * a + 0.25 * (torch.roll(a, 1, ax) + torch.roll(a, -1, ax))
        return a
    u0 = smooth(torch.randn((N, N, N + 1), device=dev, dtype=torch.float64, generator=g)); u0[:, :, -1] = u0[:, :, 0]
    v0 = smooth(torch.randn((N, N + 1, N), device=dev, dtype=torch.float64, generator=g)); v0[:, -1] = v0[:, 0]
    w0 = smooth(torch.randn((N + 1, N, N), device=dev, dtype=torch.float64, generator=g)); w0[0] = 0; w0[-1] = 0
    proj = mp.MacProjector((N, N, N), (1.0 / N,) * 3, (0, 0, 1), (0, 0, 1))

    def rel(a, b):
        return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))

    def run(scale_u=1.0, scale_b=1.0):
        cb = [scale_b * bx, scale_b * by, scale_b * bz]
        torch.cuda.synchronize()
        proj.updateCoeffs(cb)
        u, v, w = scale_u * u0, scale_u * v0, scale_u * w0
        phi = torch.zeros((N, N, N), device=dev, dtype=torch.float64)
        torch.cuda.synchronize()   # the handle runs on its own stream: torch's fills must have landed
        st = proj.project(u, v, w, 1e-11, 1e-14, mac_phi=phi)
        bnorm = max(st.rhsnorm, st.resnorm0)
        assert st.status == 0 and st.resnorm <= 1e-11 * bnorm
        div = (u[:, :, 1:] - u[:, :, :-1]) * N + (v[:, 1:] - v[:, :-1]) * N + (w[1:] - w[:-1]) * N
        assert float((div - div.mean()).abs().max()) <= 2e-11 * bnorm
        return u, v, w, phi - phi.mean(), st.iters

    u1, v1, w1, p1, it1 = run()
    assert it1 <= 12
    u2, v2, w2, p2, _ = run(scale_u=-3.0)
    assert rel(u2, -3.0 * u1) < 1e-9 and rel(w2, -3.0 * w1) < 1e-9 and rel(p2, -3.0 * p1) < 1e-9
    u3, v3, w3, p3, _ = run(scale_b=4.0)
    assert rel(u3, u1) < 1e-9 and rel(v3, v1) < 1e-9 and rel(4.0 * p3, p1) < 1e-9
    proj.close()


@pytest.mark.parametrize("case", [CASES[1], CASES[2], CASES[4]], ids=["rt_walls_z", "channel_inflow_outflow", "odd_periodic_bottom"])
def test_mac_top_level_forms_agree(case, monkeypatch):
    """the V-cycle's finest level in direct form (default: relax (sol, rhs) in place, no sol += cor, norm-only top residual) against MLMG's
    correction form (B200MAC_TOP_DIRECT=0): same iteration count, same phi and velocities to the solver tolerance, both equal to the oracle"""
    name, n, dx, bclo, bchi = case
    out = []
    for form in ("1", "0"):
        monkeypatch.setenv("B200MAC_TOP_DIRECT", form)
        mg, proj, rng = _make(case, True, seed=7)
        u, v, w = rng.standard_normal((n[2], n[1], n[0] + 1)), rng.standard_normal((n[2], n[1] + 1, n[0])), rng.standard_normal((n[2] + 1, n[1], n[0]))
        vel = [u, v, w]
        for d, ax in ((0, 2), (1, 1), (2, 0)):
            sl = lambda s: tuple(s if a == ax else slice(None) for a in range(3))
            if bclo[d] == 0:
                vel[d][sl(n[d])] = vel[d][sl(0)]
        if form == "1":
            ou, ov, ow = u.copy(), v.copy(), w.copy()
            ref = mo.project(mg.p, ou, ov, ow, mg.lv[0].b, 1e-11, 1e-14)
        phi = np.zeros((n[2], n[1], n[0]))
        st = proj.project(u, v, w, 1e-11, 1e-14, mac_phi=phi)
        assert st.status == 0 and abs(st.iters - ref["stats"]["iters"]) <= 1
        out.append((u, v, w, phi - phi.mean(), st.iters))
        proj.close()
    assert out[0][4] == out[1][4]
    for a, b in zip(out[0][:4], out[1][:4]):
        assert rel(a, b) < 1e-9
    assert rel(out[0][0], ou) < 1e-9 and rel(out[0][2], ow) < 1e-9 and rel(out[0][3], ref["phi"] - ref["phi"].mean()) < 1e-9
