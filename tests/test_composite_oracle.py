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
    assert len(GOLD) >= 4


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
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
