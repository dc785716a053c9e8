"""CPU tests of the EB (cut cell) nodal projection oracle (oracle/eb_oracle.py) and of the synthetic EB geometry
(incflo_b200/eb_geometry.py): golden fixtures from direct quadrature over the cut-cell polyhedra + sparse direct solve
(tests/golden/make_golden_eb.py), and algebraic identities."""
import itertools
import os

import numpy as np
import pytest
import scipy.sparse as sp

from incflo_b200 import eb_geometry as eg
from oracle import eb_oracle as eo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["eb_channel_cylinder", "eb_sphere_periodic_var", "eb_ramp_walls_var", "eb_cylinder_ebflow"]


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())


def load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    p = eo.Params(tuple(g["n"]), tuple(g["dx"]), g["bclo"], g["bchi"])
    sigma = g["sigma"] if g["sigma"].ndim else float(g["sigma"])
    ebv = g["eb_vel"] if g["eb_vel"].size else None
    return g, p, sigma, ebv


def test_cut_cell_axis_aligned_closed_form():
    # fluid = {x <= 0.2}: V = 0.7, int x = (0.2^2 - 0.5^2) / 2, int x^2 = (0.2^3 + 0.5^3) / 3, int x^2 y^2 = that / 12; EB face area 1 at x = 0.2
    V, S, B = eg.cut_cell((1.0, 0.0, 0.0), 0.2)
    assert abs(V - 0.7) < 1e-14
    assert abs(S[0] - (0.04 - 0.25) / 2) < 1e-14 and abs(S[1]) < 1e-14 and abs(S[2]) < 1e-14
    assert abs(S[3] - (0.008 + 0.125) / 3) < 1e-14 and abs(S[4] - 0.7 / 12) < 1e-14
    assert abs(S[15] - (0.008 + 0.125) / 36) < 1e-14 and abs(S[17] - 0.7 / 144) < 1e-14
    assert abs(B[0] - 1.0) < 1e-14 and abs(B[1] - 0.2) < 1e-14 and abs(B[4]) < 1e-14


def test_cut_cell_divergence_theorem():
    # int_F d/dx (x) = V = sum over the faces of int x n_x dA: the EB face contributes B_x n_x, the cell faces their open area * (+-1/2);
    # checked for a generic plane through int_F div(x, y, z) = 3 V = int_{dF} x . n
    rng = np.random.default_rng(3)
    for _ in range(5):
        n = rng.standard_normal(3)
        n /= np.linalg.norm(n)
        off = rng.uniform(-0.3, 0.3)
        V, S, B = eg.cut_cell(n, off)
        tets, tris = eg.cut_cell_simplices(n, off)
        # boundary of the polyhedron: the hull facets; x . n_out is constant on a facet = distance of its plane from the origin
        from scipy.spatial import ConvexHull
        pts = np.unique(np.round(tets[:, 1:, :].reshape(-1, 3), 13), axis=0)
        hull = ConvexHull(pts)
        flux = 0.0
        for s, eq in zip(hull.simplices, hull.equations):
            v = pts[s]
            area = 0.5 * np.linalg.norm(np.cross(v[1] - v[0], v[2] - v[0]))
            flux += area * (-eq[3])
        assert abs(flux - 3.0 * V) < 1e-12
        assert abs(B[0] * off - (n[0] * B[1] + n[1] * B[2] + n[2] * B[3])) < 1e-13   # the face centroid lies on the plane


def test_sphere_volume_converges():
    # tangent planes: second-order accurate geometry
    n, h = (24, 24, 24), 1.0 / 24
    g = eg.sphere(n, h, 0.3, (0.5, 0.5, 0.5))
    covered = (1.0 - g.vfrac).sum() * h ** 3
    assert abs(covered - 4.0 / 3.0 * np.pi * 0.3 ** 3) / covered < 1e-2
    area = g.barea.sum() * h ** 2
    assert abs(area - 4.0 * np.pi * 0.3 ** 2) / area < 1e-2


def test_regular_cells_give_the_27_point_operator():
    # SURVEY A.3, isotropic, sigma = 1: centre -8/(3 h^2), faces 0, edges 1/(6 h^2), corners 1/(12 h^2)
    n, h = (8, 8, 8), 0.125
    p = eo.Params(n, (h,) * 3, (0, 0, 0), (0, 0, 0))
    g = eg.EBGeometry(n)
    L = eo.build_level0(p, 1.0, g.vfrac, g.intg)
    assert np.allclose(L.st[13] * h * h, -8.0 / 3.0, atol=1e-12)
    for m, off in enumerate(eo.FWD):
        nz = sum(abs(o) for o in off)
        want = {1: 0.0, 2: 1.0 / 6.0, 3: 1.0 / 12.0}[nz]
        assert np.allclose(L.st[m] * h * h, want, atol=1e-12)


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_golden(name):
    g, p, sigma, ebv = load(name)
    r = eo.project(p, g["vel"], sigma, g["vfrac"], g["intg"], 1e-12, 1e-15, ebv, g["bnorm"], g["bintg"])
    act = r["mg"].lv[0].active
    a, b = r["phi"].copy(), g["phi"].copy()
    if p.singular:
        a[act] -= a[act].mean()
        b[act] -= b[act].mean()
    assert rel(r["info"]["rhs"], g["rhs"]) < 1e-13
    assert rel(a, b) < 1e-9 and rel(r["vel"], g["vel_new"]) < 1e-9 and rel(r["gphi"], g["gphi"]) < 1e-9
    assert r["info"]["iters"] <= 12


@pytest.mark.parametrize("name", FIXTURES[:3])
def test_stencil_identities(name):
    g, p, sigma, _ = load(name)
    mg = eo.MG(p, sigma, g["vfrac"], g["intg"])
    rng = np.random.default_rng(1)
    for li, L in enumerate(mg.lv):
        x = rng.standard_normal(L.shape)
        # stencil form == matrix form; symmetric; negative semi-definite; inactive rows and columns vanish
        A = L.matrix()
        assert np.abs(L.apply_stencil(x) - L.apply(x)).max() < 1e-9 * np.abs(A).max()
        assert abs(A - A.T).max() < 1e-12 * np.abs(A).max()
        xa = np.where(L.active, x, 0.0)
        assert np.vdot(xa, L.apply(xa)) < 0.0
        assert np.abs(L.apply(np.where(L.active, 0.0, 1.0))).max() == 0.0
        if p.singular:   # constants on the active nodes are the null space
            assert np.abs(L.apply(L.active.astype(float))).max() < 1e-10 * np.abs(A).max()
        if li + 1 < len(mg.lv):   # Galerkin: A_c = (1/8) P^T A P with an explicitly assembled trilinear P
            C = mg.lv[li + 1]
            P = trilinear_matrix(L, C)
            Mc = sp.diags((~C.dmask).ravel().astype(float))
            Ac = Mc @ (P.T @ A @ P) @ Mc * 0.125
            assert abs(Ac - C.matrix()).max() < 1e-12 * np.abs(A).max()
            xc = rng.standard_normal(C.shape)
            assert np.abs(eo.interp_add(L, C, np.zeros(L.shape), xc) - np.where(L.active, (P @ xc.ravel()).reshape(L.shape), 0.0)).max() < 1e-13
            r = np.where(L.active, x, 0.0)
            assert np.abs(eo.restrict(L, C, r) - np.where(C.active, (P.T @ r.ravel()).reshape(C.shape) * 0.125, 0.0)).max() < 1e-13


def trilinear_matrix(F, C):
    """P (fine nodes x coarse nodes), assembled entry by entry"""
    nf, nc = F.nn, C.nn
    rows, cols, vals = [], [], []
    fi = np.arange(int(np.prod(F.shape))).reshape(F.shape)
    ci = np.arange(int(np.prod(C.shape))).reshape(C.shape)
    for k, j, i in itertools.product(range(nf[2]), range(nf[1]), range(nf[0])):
        opts = []
        for d, q in enumerate((i, j, k)):
            if q % 2 == 0:
                opts.append([(q // 2, 1.0)])
            else:
                hi = (q + 1) // 2
                if hi >= nc[d]:
                    hi = 0 if F.per[d] else None
                opts.append([(q // 2, 0.5)] + ([(hi, 0.5)] if hi is not None else []))
        for (ic, wi), (jc, wj), (kc, wk) in itertools.product(*opts):
            rows.append(fi[k, j, i]); cols.append(ci[kc, jc, ic]); vals.append(wi * wj * wk)
    return sp.csr_matrix((vals, (rows, cols)), shape=(fi.size, ci.size))


def test_divergence_is_minus_gradient_transpose():
    # <D u, phi> = - sum_c V_c u_c . (cell average of grad phi)   (periodic box, no boundary terms)
    g, p, sigma, _ = load("eb_sphere_periodic_var")
    L0 = eo.build_level0(p, sigma, g["vfrac"], g["intg"])
    rng = np.random.default_rng(5)
    phi = np.where(L0.active, rng.standard_normal(L0.shape), 0.0)
    vel = rng.standard_normal(g["vel"].shape)
    rhs = eo.compute_rhs(p, L0, vel, g["vfrac"], g["intg"])
    gr = eo.gradient(p, L0, phi, g["vfrac"], g["intg"])
    lhs = np.vdot(rhs, phi)
    rhs2 = -np.sum(g["vfrac"][None] * vel[:, 1:-1, 1:-1, 1:-1] * gr)
    assert abs(lhs - rhs2) < 1e-10 * abs(lhs)


def test_odd_periodic_level_converges():
    n, h = (12, 12, 12), 1.0 / 12
    geom = eg.sphere(n, h, 0.21, (0.5, 0.45, 0.55), small_vfrac=1e-3)
    p = eo.Params(n, (h,) * 3, (0, 0, 0), (0, 0, 0))
    rng = np.random.default_rng(2)
    vel = rng.standard_normal((3, 14, 14, 14)) * (np.pad(geom.vfrac, 1) > 0)
    r = eo.project(p, vel, 1.0, geom.vfrac, geom.intg, 1e-10, 1e-14)
    assert [L.n for L in r["mg"].lv] == [(12, 12, 12), (6, 6, 6), (3, 3, 3)] and r["mg"].lv[2].odd_periodic
    assert r["info"]["iters"] <= 12


def test_uniform_flow_past_a_cylinder_converges_at_second_order():
    """Projecting u = (1, 0, 0) in the fluid around a cylinder (periodic box) must give the potential flow around it: the natural boundary
    condition of the fluid-integrated operator is no flux through the body.  No closed form in a periodic box, so: self-convergence of the
    projected velocity under refinement by 2 (volume-weighted 2 x 2 averages of the finer solution) -- second order away from the
    body (ratio ~ 4), between first and second order over all fluid cells"""
    def solve(N):
        n, h = (N, N, N // 8), 1.0 / N
        geom = eg.cylinder(n, h, 0.1500001, (0.5, 0.5, 0.0), direction=2, small_vfrac=1e-3)
        p = eo.Params(n, (h,) * 3, (0, 0, 0), (0, 0, 0))
        vel = np.zeros((3, n[2] + 2, N + 2, N + 2))
        vel[0, 1:-1, 1:-1, 1:-1] = (geom.vfrac > 0)
        r = eo.project(p, vel, 1.0, geom.vfrac, geom.intg, 1e-12, 1e-15)
        assert np.abs(r["vel"][2]).max() < 1e-10          # the problem is two-dimensional
        return geom.vfrac[0], r["vel"][:2, 0]

    def restrict(u, V):
        parts = [(slice(None), slice(a, None, 2), slice(b, None, 2)) for a in (0, 1) for b in (0, 1)]
        den = sum(V[s[1:]] for s in parts)
        num = sum(u[s] * V[s[1:]] for s in parts)
        return np.where(den > 0, num / np.where(den > 0, den, 1.0), 0.0)

    sol = {N: solve(N) for N in (32, 64, 128)}
    err = {}
    for a, b in ((32, 64), (64, 128)):
        Va, ua = sol[a]
        ub = restrict(sol[b][1], sol[b][0])
        x = (np.arange(a) + 0.5) / a
        X, Y = np.meshgrid(x, x, indexing="xy")
        far = (X - 0.5) ** 2 + (Y - 0.5) ** 2 > 0.2 ** 2
        err[a] = (np.sqrt(np.mean((ub - ua)[:, far] ** 2)), np.sqrt(np.mean((ub - ua)[:, Va > 0] ** 2)))
    assert err[32][0] / err[64][0] > 3.3 and err[64][0] < 3e-4
    assert err[32][1] / err[64][1] > 2.2
    # the flow is deflected around the body: faster than the free stream above / below it, slower in front of it
    V, u = sol[128]
    assert u[0, 64 + 26, 64] > 1.2 and u[0, 64, 64 - 26] < 0.8
