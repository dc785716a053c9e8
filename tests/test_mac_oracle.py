"""MAC projection oracle (oracle/mac_oracle.py, numpy) against the independent SciPy direct solves of
tests/golden/make_golden_mac.py and against algebraic identities (CPU only)."""
import glob
import os

import numpy as np
import pytest

from oracle import mac_oracle as mo

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mac_*.npz")))


def load_mac(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    for k in ("n", "bclo", "bchi"):
        d[k] = tuple(int(x) for x in d[k])
    d["dx"] = tuple(float(x) for x in d["dx"])
    return d


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def test_mac_fixtures_present():
    assert len(GOLD) == 4


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_mac_oracle_reproduces_golden(path):
    g = load_mac(path)
    p = mo.Params(g["n"], g["dx"], g["bclo"], g["bchi"])
    u, v, w = g["u_in"].copy(), g["v_in"].copy(), g["w_in"].copy()
    r = mo.project(p, u, v, w, [g["bx"], g["by"], g["bz"]], 1e-12, 0.0)
    st = r["stats"]
    assert st["status"] == 0 and st["iters"] <= 30
    singular = all(b != mo.DIR for b in g["bclo"] + g["bchi"])
    a, b = (r["phi"] - r["phi"].mean(), g["phi"] - g["phi"].mean()) if singular else (r["phi"], g["phi"])
    assert rel(a, b) < 1e-9
    assert rel(u, g["u_out"]) < 1e-9 and rel(v, g["v_out"]) < 1e-9 and rel(w, g["w_out"]) < 1e-9


def test_operator_is_symmetric_and_annihilates_constants():
    rng = np.random.default_rng(0)
    n = (8, 6, 4)
    b = [rng.uniform(0.5, 2, size=(4, 6, 9)), rng.uniform(0.5, 2, size=(4, 7, 8)), rng.uniform(0.5, 2, size=(5, 6, 8))]
    for bclo, bchi in (((0, 1, 1), (0, 1, 1)), ((2, 1, 0), (1, 2, 0))):
        if bclo[0] == 0:
            b[0][:, :, -1] = b[0][:, :, 0]
        if bclo[2] == 0:
            b[2][-1] = b[2][0]
        mg = mo.MG(mo.Params(n, (0.1, 0.2, 0.3), bclo, bchi, max_coarsening_level=0), b)
        x, y = rng.standard_normal((4, 6, 8)), rng.standard_normal((4, 6, 8))
        if mg.singular:   # (the maxorder-3 Dirichlet stencil 3 phi_0 - phi_1 / 3 makes the operator non-symmetric, as in AMReX)
            assert abs((y * mg.adotx(0, x)).sum() - (x * mg.adotx(0, y)).sum()) < 1e-9 * abs((y * mg.adotx(0, x)).sum())
            assert np.abs(mg.adotx(0, np.ones((4, 6, 8)))).max() < 1e-10
        else:             # a quadratic that vanishes on the Dirichlet face x = 0 is differentiated exactly there
            xc = (np.arange(8) + 0.5) * 0.1
            q = np.broadcast_to(xc * xc + 2.0 * xc, (4, 6, 8)).copy()
            mgc = mo.MG(mo.Params(n, (0.1, 0.2, 0.3), bclo, bchi, max_coarsening_level=0), 1.0)
            assert np.allclose(mgc.adotx(0, q)[:, 1:-1, :4], -2.0)


def test_smoother_reduces_the_residual_and_bottom_solver_converges():
    rng = np.random.default_rng(1)
    n = (16, 16, 16)
    p = mo.Params(n, (1 / 16,) * 3, (0, 0, 1), (0, 0, 1))
    mg = mo.MG(p, 0.5)
    rhs = rng.standard_normal((16, 16, 16)); rhs -= rhs.mean()
    phi = np.zeros_like(rhs)
    r0 = np.abs(mg.residual(0, phi, rhs)).max()
    mg.smooth(0, phi, rhs, 4)
    assert np.abs(mg.residual(0, phi, rhs)).max() < 0.7 * r0
    lev = len(mg.lv) - 1
    nb = mg.lv[lev].n
    b = rng.standard_normal((nb[2], nb[1], nb[0])); b -= b.mean()
    x = mg.bottom_solve(np.zeros_like(b), b)
    assert np.abs(mg.residual(lev, x, b)).max() <= 1e-4 * np.abs(b).max() * 1.01


def test_projection_is_idempotent():
    g = load_mac(GOLD[2])
    p = mo.Params(g["n"], g["dx"], g["bclo"], g["bchi"])
    u, v, w = g["u_out"].copy(), g["v_out"].copy(), g["w_out"].copy()
    r = mo.project(p, u, v, w, [g["bx"], g["by"], g["bz"]], 1e-11, 1e-9)
    assert r["stats"]["iters"] == 0 and rel(u, g["u_out"]) < 1e-12
