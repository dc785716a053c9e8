"""Composite (two AMR level, one fine box at ratio 2) nodal projection, CPU side: the oracle
(oracle/composite.py) against the golden fixtures tests/golden/composite/*.npz -- an independent SciPy
direct solve of the composite Q1 finite-element system with hanging-node constraints
(tests/golden/make_golden_composite.py).  Tolerance 1e-9 relative L2 (north_star) at rtol 1e-12."""
import glob
import os

import numpy as np
import pytest

from helpers import oracle_params, rel_l2

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "composite", "*.npz")))
GOLD_DIR = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "composite_dirichlet", "*.npz")))
TOL = 1e-9


def load(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    for k in ("n0", "bclo", "bchi", "clo", "chi"):
        g[k] = tuple(int(x) for x in g[k])
    g["dx0"] = tuple(float(x) for x in g["dx0"])
    g["ng0"], g["ng1"], g["var"] = int(g["ng0"]), int(g["ng1"]), bool(g["var"])
    return g


def to_full(p, bclo):
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        if bclo[d] == 0:
            p = np.concatenate([p, np.take(p, [0], axis=ax)], axis=ax)
    return p


def fine_per(g):
    """(x, y, z) BC codes with 0 where the fine box spans a periodic direction (unique-node layout there)"""
    return tuple(0 if (g["bclo"][d] == 0 and g["clo"][d] == 0 and g["chi"][d] == g["n0"][d] - 1) else 1 for d in range(3))


def check(g, vel0, vel1, phi0_full, phi1, gphi0, gphi1, tol=TOL):
    """phi is defined up to one constant common to both levels (all cases are singular)"""
    n0 = g["n0"]; nf = tuple(2 * (h - l + 1) for l, h in zip(g["clo"], g["chi"]))
    c = phi1.mean() - g["phi1"].mean()
    assert rel_l2(phi1 - c, g["phi1"]) < tol
    assert rel_l2(phi0_full - c, g["phi0"]) < tol
    assert rel_l2(gphi1, g["gphi1"]) < tol and rel_l2(gphi0, g["gphi0"]) < tol
    a, b = g["ng0"], g["ng1"]
    i0 = (slice(None), slice(a, a + n0[2]), slice(a, a + n0[1]), slice(a, a + n0[0]))
    i1 = (slice(None), slice(b, b + nf[2]), slice(b, b + nf[1]), slice(b, b + nf[0]))
    assert rel_l2(vel0[i0], g["vel0_out"][i0]) < tol and rel_l2(vel1[i1], g["vel1_out"][i1]) < tol


def test_fixtures_present():
    assert len(GOLD) >= 4 and len(GOLD_DIR) >= 2


@pytest.mark.parametrize("path", GOLD + GOLD_DIR, ids=[os.path.basename(p)[:-4] for p in GOLD + GOLD_DIR])
def test_composite_oracle_reproduces_golden(path, oracle):
    from oracle import composite as oc
    g = load(path)
    p0 = oracle_params(g["n0"], g["dx0"], g["bclo"], g["bchi"])
    cp = oc.CompositeProjector(p0, g["clo"], g["chi"])
    v0, v1 = g["vel0_in"].copy(), g["vel1_in"].copy()
    r = cp.project(v0, g["ng0"], v1, g["ng1"], g["sigma0"] if g["var"] else None, g["sigma1"] if g["var"] else None,
                   float(g["sigma0"].flat[0]), rtol=1e-12, atol=0.0)
    assert r["status"] == 0 and r["iters"] <= 20
    check(g, v0, v1, to_full(r["phi0"], g["bclo"]), to_full(r["phi1"], fine_per(g)), r["gphi0"], r["gphi1"])


def test_composite_converges_and_removes_a_gradient(oracle):
    """u = grad(psi) sampled on both levels (psi smooth): the composite projection must converge to rtol and
    remove the gradient to O(h^2) on both levels (an approximate projection is not idempotent -- L != D sigma G,
    SURVEY 0.2 -- so "projecting twice changes nothing" is NOT a property of this operator)."""
    from oracle import composite as oc
    N = 16
    p0 = oracle_params((N, N, N), (1.0 / N,) * 3, (0, 0, 1), (0, 0, 1))
    tp = 2 * np.pi

    def grad_psi(n, h, off):   # psi = cos(2 pi x) cos(2 pi y) cos(pi z): periodic in x, y, zero normal derivative at z = 0, 1
        z, y, x = np.meshgrid(*[(np.arange(m) + 0.5) * h + o for m, o in zip(n[::-1], off[::-1])], indexing="ij")
        return np.stack([-tp * np.sin(tp * x) * np.cos(tp * y) * np.cos(np.pi * z),
                         -tp * np.cos(tp * x) * np.sin(tp * y) * np.cos(np.pi * z),
                         -np.pi * np.cos(tp * x) * np.cos(tp * y) * np.sin(np.pi * z)])
    clo, chi = (4, 4, 4), (11, 11, 11)
    vel0 = np.zeros((3, N + 2, N + 2, N + 2)); vel0[:, 1:-1, 1:-1, 1:-1] = grad_psi((N, N, N), 1.0 / N, (0, 0, 0))
    vel1 = np.zeros((3, 18, 18, 18)); vel1[:, 1:-1, 1:-1, 1:-1] = grad_psi((16, 16, 16), 0.5 / N, (4.0 / N,) * 3)
    u0, u1 = np.abs(vel0).max(), np.abs(vel1).max()
    cp = oc.CompositeProjector(p0, clo, chi)
    r = cp.project(vel0, 1, vel1, 1, const_sigma=1.0, rtol=1e-11)
    assert r["status"] == 0 and r["iters"] <= 20
    assert r["resnorm"] <= 1e-11 * max(r["rhsnorm"], r["resnorm0"])
    assert np.abs(vel1[:, 1:-1, 1:-1, 1:-1]).max() < 0.05 * u1      # O(h^2) remainder on the fine level
    assert np.abs(vel0[:, 1:-1, 1:-1, 1:-1]).max() < 0.12 * u0      # coarser level: 4x larger


@pytest.mark.parametrize("case", ["interior_box", "periodic_span_slab", "corner_on_walls"])
def test_composite_operator_equals_assembled_finite_element_matrix(case, oracle):
    """Column by column: the oracle's composite residual operator (reflux through the reflected fine box and
    the sigma-masked coarse level) against the Q1 finite-element matrix A of the composite mesh with hanging-node
    constraints, assembled independently in tests/golden/make_golden_composite.py -- entry-wise, not just for
    one right-hand side.  Rows are compared in natural (element-sum) scaling: the oracle's equations are the
    element sums divided by the cell volume and by the node weight 1/2 per wall face.
    Fine rows must be identical.  A coarse row I on the interface is the FE row PLUS the full-weighting share
    of the residuals of the fine UNKNOWNS next to the interface (restriction does not stop at the hanging
    nodes; AMReX's reflux restricts the fine residual the same way): op = [[1, W], [0, 1]] A with W the trilinear
    weights -- an equivalent system (same solutions), which is why the golden solutions agree to 1e-13."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_composite as mg
    from oracle import composite as oc
    PER, NEU = 0, 1
    if case == "interior_box":
        n0, bc, clo, chi = (8, 8, 6), (PER, NEU, NEU), (2, 3, 1), (5, 5, 3)
    elif case == "periodic_span_slab":
        n0, bc, clo, chi = (6, 6, 8), (PER, PER, NEU), (0, 0, 2), (5, 5, 4)
    else:
        n0, bc, clo, chi = (6, 6, 6), (NEU, NEU, NEU), (0, 2, 3), (2, 4, 5)
    dx0 = (1.0 / 8, 1.0 / 8, 1.0 / 8)
    nf = tuple(2 * (h - l + 1) for l, h in zip(clo, chi))
    rng = np.random.default_rng(3)
    s0 = rng.uniform(0.5, 2.0, size=n0[::-1]); s1 = rng.uniform(0.5, 2.0, size=nf[::-1])
    z0 = np.zeros((3, n0[2] + 2, n0[1] + 2, n0[0] + 2)); z1 = np.zeros((3, nf[2] + 2, nf[1] + 2, nf[0] + 2))
    sysm = mg.solve(n0, dx0, bc, bc, clo, chi, z0, 1, z1, 1, s0, s1, assemble_only=True)
    A = sysm["A"].toarray()
    cp = oc.CompositeProjector(oracle.make_params(n0, dx0, bc, bc), clo, chi)
    cp.setup(s0, s1)
    shape0, shape1 = cp.mg0.node_shape(0), cp.mgD.node_shape(0)
    N0 = int(np.prod(shape0))
    act0 = sysm["act"][:N0].reshape(shape0)                 # coarse nodes that are unknowns
    int1 = np.zeros(int(np.prod(shape1)), dtype=bool); int1[sysm["int_ids"]] = True
    int1 = int1.reshape(shape1)                              # fine nodes that are unknowns (not hanging)
    w0 = cp.mg0.dot_weights(0); w1 = cp.mgN.dot_weights(0)   # 1/2 per wall face
    idx0, idx1 = np.argwhere(act0), np.argwhere(int1)
    n_act0 = len(idx0); n_unk = n_act0 + len(idx1)
    assert A.shape == (n_unk, n_unk)
    op = np.zeros_like(A)
    for j in range(n_unk):
        sol0 = np.zeros(shape0); sol1 = np.zeros(shape1)
        if j < n_act0:
            sol0[tuple(idx0[j])] = 1.0
        else:
            sol1[tuple(idx1[j - n_act0])] = 1.0
        cp.fill_hanging(sol0, sol1)
        r0, r1 = cp.composite_residual(sol0, sol1)           # = -A_composite x in the oracle's row scaling
        op[:n_act0, j] = -(r0 * w0)[act0] * sysm["H3"]
        op[n_act0:, j] = -(r1 * w1)[int1] * sysm["h3"]
    # W: trilinear weight of the fine unknown j in the coarse basis function of the interface node I
    fid = -np.ones(shape1, dtype=np.int64); fid[int1] = np.arange(len(idx1))
    W = np.zeros((n_act0, len(idx1)))
    lo = (clo[2], clo[1], clo[0]); span = cp.span[::-1]; cf = (list(zip(cp.cf_lo, cp.cf_hi)))[::-1]
    for I, (k, j, i) in enumerate(idx0):
        c = (k, j, i)
        inside = all(span[a] or lo[a] <= c[a] <= lo[a] + cp.nb[2 - a] for a in range(3))
        on_cf = any((cf[a][0] and c[a] == lo[a]) or (cf[a][1] and c[a] == lo[a] + cp.nb[2 - a]) for a in range(3))
        if not (inside and on_cf):
            continue
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    f = [2 * (c[a] - lo[a]) + d for a, d in enumerate((dz, dy, dx))]
                    ok = True
                    for a in range(3):
                        if span[a]:
                            f[a] %= shape1[a]
                        elif not 0 <= f[a] < shape1[a]:
                            ok = False
                    if ok and fid[tuple(f)] >= 0:
                        W[I, fid[tuple(f)]] += (1 - abs(dz) / 2) * (1 - abs(dy) / 2) * (1 - abs(dx) / 2)
    expected = A.copy()
    expected[:n_act0] += W @ A[n_act0:]
    scale = np.abs(A).max()
    assert np.abs(op[n_act0:] - A[n_act0:]).max() < 1e-12 * scale      # fine rows: the FE rows themselves
    assert np.abs(op - expected).max() < 1e-12 * scale                  # coarse rows: FE row + restricted fine rows
    assert np.abs(A - A.T).max() < 1e-12 * scale and np.abs(op.sum(axis=1)).max() < 1e-11 * scale


def test_composite_solution_converges_at_second_order(oracle):
    """u = grad(psi) on both levels: the composite phi must approach psi (up to a constant) at O(h^2) on the fine
    AND on the coarse level when the whole hierarchy is refined (pins sign and scaling conventions of the
    composite rhs, reflux and interpolation against an analytic solution, like SURVEY 8(c).4 does for one level)"""
    from oracle import composite as oc
    tp = 2 * np.pi

    def psi(x, y, z):
        return np.cos(tp * x) * np.cos(tp * y) * np.cos(np.pi * z)

    def grad_psi(n, h, off):
        z, y, x = np.meshgrid(*[(np.arange(m) + 0.5) * h + o for m, o in zip(n[::-1], off[::-1])], indexing="ij")
        return np.stack([-tp * np.sin(tp * x) * np.cos(tp * y) * np.cos(np.pi * z),
                         -tp * np.cos(tp * x) * np.sin(tp * y) * np.cos(np.pi * z),
                         -np.pi * np.cos(tp * x) * np.cos(tp * y) * np.sin(np.pi * z)])
    errs = []
    for N in (16, 32):
        p0 = oracle_params((N, N, N), (1.0 / N,) * 3, (0, 0, 1), (0, 0, 1))
        clo, chi = (N // 4,) * 3, (3 * N // 4 - 1,) * 3
        vel0 = np.zeros((3, N + 2, N + 2, N + 2)); vel0[:, 1:-1, 1:-1, 1:-1] = grad_psi((N, N, N), 1.0 / N, (0, 0, 0))
        vel1 = np.zeros((3, N + 2, N + 2, N + 2)); vel1[:, 1:-1, 1:-1, 1:-1] = grad_psi((N, N, N), 0.5 / N, (clo[0] / N,) * 3)
        r = oc.CompositeProjector(p0, clo, chi).project(vel0, 1, vel1, 1, const_sigma=1.0, rtol=1e-12, atol=0.0)
        assert r["status"] == 0
        c0 = np.arange(N + 1) / N
        Z, Y, X = np.meshgrid(c0, c0, c0, indexing="ij")
        f0 = clo[0] / N + np.arange(N + 1) * 0.5 / N
        Zf, Yf, Xf = np.meshgrid(f0, f0, f0, indexing="ij")
        phi0 = to_full(r["phi0"], (0, 0, 1))
        c = (phi0 - psi(X, Y, Z)).mean()
        errs.append((np.abs(phi0 - c - psi(X, Y, Z)).max(), np.abs(r["phi1"] - c - psi(Xf, Yf, Zf)).max()))
    assert errs[1][0] < 0.3 * errs[0][0] and errs[1][1] < 0.3 * errs[0][1]     # ~ 1/4 per halving of h
    assert errs[1][0] < 5e-3 and errs[1][1] < 5e-3
