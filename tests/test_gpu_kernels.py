"""Per-kernel parity: every CUDA building block of the V-cycle, called through the C ABI
(b200np_level_set / _op / _get), against the CPU oracle on the same seeded inputs.

Tolerances (fp64): these kernels evaluate the same formulas as the oracle with a different
summation order, so results agree to a few ulp of the largest intermediate; we require
1e-12 relative to the max-norm of the result (stated per test).
"""
import numpy as np
import pytest

from helpers import BC_CASES, TILE, oracle_params, random_sigma

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _setup(case, var, oracle, seed=0):
    from incflo_b200 import nodal_projector as npj
    name, n, dx, bclo, bchi = case
    rng = np.random.default_rng(seed)
    sigma = random_sigma(n, rng) if var else None
    csig = 0.7
    p = oracle_params(n, dx, bclo, bchi)
    mg = oracle.MG(p, sigma, csig)
    vel = np.zeros((3, n[2] + 2, n[1] + 2, n[0] + 2))
    proj = npj.NodalProjector(vel, sigma, csig, dict(n_cell=n, dx=dx, is_periodic=[b == 0 for b in bclo]), ng=1,
                              opts=npj.nodal_proj_opts(tile=TILE))
    proj.setDomainBC(bclo, bchi)
    proj.set_sigma(sigma, csig)
    assert proj.nlevels() == mg.nlev
    return mg, proj, rng


def _masked_random(mg, lev, rng):
    x = rng.standard_normal(mg.node_shape(lev))
    w = mg.dot_weights(lev)
    x[w == 0] = 0.0
    return x


def _close(a, b, tol=RTOL):
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err < tol, f"max rel err {err:.3e}"


@pytest.mark.parametrize("var", [False, True])
@pytest.mark.parametrize("case", BC_CASES, ids=[c[0] for c in BC_CASES])
def test_coarsen_sigma_and_dims(case, var, oracle):
    mg, proj, rng = _setup(case, var, oracle)
    for lev in range(mg.nlev):
        assert proj.level_dims(lev) == mg.dims(lev)
        if var:
            from incflo_b200.nodal_projector import A_SIGMA
            _close(proj.level_get(lev, A_SIGMA), mg.sigma(lev), 1e-15)  # same additions, exact up to order


@pytest.mark.parametrize("var", [False, True])
@pytest.mark.parametrize("case", BC_CASES, ids=[c[0] for c in BC_CASES])
def test_residual(case, var, oracle):
    from incflo_b200.nodal_projector import A_COR, A_RES, A_RESCOR, OP_RESIDUAL
    mg, proj, rng = _setup(case, var, oracle)
    for lev in range(mg.nlev):
        phi = _masked_random(mg, lev, rng)
        rhs = _masked_random(mg, lev, rng)
        proj.level_set(lev, A_COR, phi); proj.level_set(lev, A_RES, rhs)
        proj.level_op(lev, OP_RESIDUAL)
        got = proj.level_get(lev, A_RESCOR)
        _close(got, mg.residual(lev, phi, rhs))


@pytest.mark.parametrize("nsweeps", [1, 4])
@pytest.mark.parametrize("var", [False, True])
@pytest.mark.parametrize("case", BC_CASES, ids=[c[0] for c in BC_CASES])
def test_smoother_sweeps(case, var, nsweeps, oracle):
    """k_smooth_tile == oracle Gauss-Seidel in box mode (tile 64x16x16, plane-4-colour order,
    previous-sweep values outside the tile), sweep by sweep."""
    from incflo_b200.nodal_projector import A_COR, A_RES, OP_SMOOTH
    mg, proj, rng = _setup(case, var, oracle)
    for lev in range(mg.nlev):
        phi = _masked_random(mg, lev, rng)
        rhs = _masked_random(mg, lev, rng)
        proj.level_set(lev, A_COR, phi); proj.level_set(lev, A_RES, rhs)
        proj.level_op(lev, OP_SMOOTH, nsweeps)
        got = proj.level_get(lev, A_COR)
        ref = mg.smooth(lev, phi.copy(), rhs, nsweeps)
        _close(got, ref, 1e-11)


@pytest.mark.parametrize("var", [False, True])
@pytest.mark.parametrize("case", BC_CASES, ids=[c[0] for c in BC_CASES])
def test_restriction(case, var, oracle):
    from incflo_b200.nodal_projector import A_RES, A_RESCOR, OP_RESTRICT
    mg, proj, rng = _setup(case, var, oracle)
    for lev in range(mg.nlev - 1):
        fine = _masked_random(mg, lev, rng)
        proj.level_set(lev, A_RESCOR, fine)
        proj.level_op(lev, OP_RESTRICT)
        _close(proj.level_get(lev + 1, A_RES), mg.restrict(lev, fine))


@pytest.mark.parametrize("var", [False, True])
@pytest.mark.parametrize("case", BC_CASES, ids=[c[0] for c in BC_CASES])
def test_interpolation(case, var, oracle):
    from incflo_b200.nodal_projector import A_COR, OP_INTERP
    mg, proj, rng = _setup(case, var, oracle)
    for lev in range(mg.nlev - 1):
        fine = _masked_random(mg, lev, rng)
        crse = _masked_random(mg, lev + 1, rng)
        proj.level_set(lev, A_COR, fine); proj.level_set(lev + 1, A_COR, crse)
        proj.level_op(lev, OP_INTERP)
        _close(proj.level_get(lev, A_COR), mg.interp_add(lev, fine.copy(), crse))


@pytest.mark.parametrize("var", [False, True])
@pytest.mark.parametrize("case", BC_CASES, ids=[c[0] for c in BC_CASES])
def test_bottom_solve(case, var, oracle):
    """single-CTA BiCGStab: same iteration (rtol 1e-4) => answers agree far below the bottom
    tolerance; compare to 1e-6 of the solution norm and check the residual criterion itself."""
    from incflo_b200.nodal_projector import A_COR, A_RES, OP_BOTTOM
    mg, proj, rng = _setup(case, var, oracle)
    lev = mg.nlev - 1
    b = _masked_random(mg, lev, rng)
    proj.level_set(lev, A_RES, b)
    proj.level_op(lev, OP_BOTTOM)
    got = proj.level_get(lev, A_COR)
    ref, its = mg.bottom_solve(b)
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 1e-6 * scale
    # residual criterion of A.10 on the mean-free rhs
    w = mg.dot_weights(lev)
    singular = all(bc != 2 for bc in case[3] + case[4])
    bb = b - (w * b).sum() / w.sum() if singular else b
    bb = np.where(w > 0, bb, 0.0)
    r = mg.residual(lev, got, bb)
    assert np.abs(r).max() <= 1.01e-4 * np.abs(bb).max()


@pytest.mark.parametrize("var", [False, True])
def test_vcycle(var, oracle):
    """one whole V-cycle (pre-smooth, residual, restriction, bottom, interpolation, post-smooth)"""
    from incflo_b200.nodal_projector import A_COR, A_RES, OP_VCYCLE
    case = BC_CASES[1]
    mg, proj, rng = _setup(case, var, oracle)
    res = _masked_random(mg, 0, rng)
    w = mg.dot_weights(0)
    res -= (w * res).sum() / w.sum()
    proj.level_set(0, A_RES, res)
    proj.level_op(0, OP_VCYCLE)
    got = proj.level_get(0, A_COR)
    # oracle V-cycle = solve with maxiter 1 from phi = 0: phi_1 = cor
    mg.params.maxiter = 1
    mg2 = oracle.MG(mg.params, mg.sigma(0), 0.7)
    phi = np.zeros_like(res)
    mg2.solve(phi, res.copy(), 1e-30, 0.0)
    _close(got, phi, 1e-8)


# ---- the kernels that run at BENCHMARK size ------------------------------------------------------
# b200np.cu routes a level with <= 148 smoother CTAs to the resident-chunk kernel (k_smooth_iso_res); every grid
# above is that small.  B200NP_RES_CTAS=0 forces the ring-slot kernel k_smooth_iso (50 % of a 256^3 solve) onto the
# same cases, and the 128^3 cases below reach it with the default routing (272 CTAs on level 0).
@pytest.mark.parametrize("zero_start", [False, True], ids=["read_cor", "zero_start"])
@pytest.mark.parametrize("nsweeps", [1, 4])
@pytest.mark.parametrize("var", [False, True])
@pytest.mark.parametrize("case", BC_CASES, ids=[c[0] for c in BC_CASES])
def test_smoother_sweeps_nonresident_kernel(case, var, nsweeps, zero_start, oracle, monkeypatch):
    """k_smooth_iso<VAR, FULL|edge, RES=false> (+ its SM_ZERO_IN path) == oracle, sweep by sweep"""
    from incflo_b200.nodal_projector import A_COR, A_RES, OP_SMOOTH, SMOOTH_ZERO_START
    monkeypatch.setenv("B200NP_RES_CTAS", "0")
    mg, proj, rng = _setup(case, var, oracle)
    for lev in range(mg.nlev):
        junk = rng.standard_normal(mg.node_shape(lev))          # zero start: cor must never be read
        phi = junk if zero_start else _masked_random(mg, lev, rng)
        rhs = _masked_random(mg, lev, rng)
        proj.level_set(lev, A_COR, phi); proj.level_set(lev, A_RES, rhs)
        proj.level_op(lev, OP_SMOOTH, nsweeps | (SMOOTH_ZERO_START if zero_start else 0))
        got = proj.level_get(lev, A_COR)
        start = np.zeros_like(phi) if zero_start else phi.copy()
        _close(got, mg.smooth(lev, start, rhs, nsweeps), 1e-11)


@pytest.mark.parametrize("zero_start", [False, True], ids=["read_cor", "zero_start"])
@pytest.mark.parametrize("var", [False, True])
def test_smoother_zero_start_resident_kernel(var, zero_start, oracle):
    """the same zero-start contract on the default routing (k_smooth_iso_res at these sizes)"""
    from incflo_b200.nodal_projector import A_COR, A_RES, OP_SMOOTH, SMOOTH_ZERO_START
    mg, proj, rng = _setup(BC_CASES[1], var, oracle)
    for lev in range(mg.nlev):
        phi = rng.standard_normal(mg.node_shape(lev)) if zero_start else _masked_random(mg, lev, rng)
        rhs = _masked_random(mg, lev, rng)
        proj.level_set(lev, A_COR, phi); proj.level_set(lev, A_RES, rhs)
        proj.level_op(lev, OP_SMOOTH, 4 | (SMOOTH_ZERO_START if zero_start else 0))
        start = np.zeros_like(phi) if zero_start else phi.copy()
        _close(proj.level_get(lev, A_COR), mg.smooth(lev, start, rhs, 4), 1e-11)


@pytest.mark.parametrize("env", [{}, {"B200NP_RES_CTAS": "0"}, {"B200NP_ZERO_START": "0"}, {"B200NP_RES_CTAS": "0", "B200NP_ZERO_START": "0"}],
                         ids=["default", "nonresident", "no_zero_start", "nonresident_no_zero_start"])
@pytest.mark.parametrize("var", [False, True])
@pytest.mark.parametrize("case", BC_CASES, ids=[c[0] for c in BC_CASES])
def test_vcycle_kernel_routing(case, var, env, oracle, monkeypatch):
    """one whole V-cycle with every kernel routing: resident / ring-slot smoother, zero-start on / off"""
    from incflo_b200.nodal_projector import A_COR, A_RES, OP_VCYCLE
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    mg, proj, rng = _setup(case, var, oracle)
    res = _masked_random(mg, 0, rng)
    w = mg.dot_weights(0)
    res -= (w * res).sum() / w.sum()
    proj.level_set(0, A_RES, res)
    proj.level_set(0, A_COR, rng.standard_normal(mg.node_shape(0)))   # junk: the V-cycle starts from cor = 0 (A.9)
    proj.level_op(0, OP_VCYCLE)
    got = proj.level_get(0, A_COR)
    mg.params.maxiter = 1
    mg2 = oracle.MG(mg.params, mg.sigma(0) if var else None, 0.7)
    phi = np.zeros_like(res)
    mg2.solve(phi, res.copy(), 1e-30, 0.0)
    _close(got, phi, 1e-8)


CASES_128 = [
    ("rt128", (128, 128, 128), (1 / 128,) * 3, (0, 0, 1), (0, 0, 1)),          # BASELINE configs[1] layout
    ("periodic128", (128, 128, 128), (1 / 128,) * 3, (0, 0, 0), (0, 0, 0)),    # configs[0] / [2] layout
    ("channel_160x64x48", (160, 64, 48), (1 / 160,) * 3, (3, 1, 0), (2, 1, 0)),  # inflow / outflow, edge tiles, multi-tile
]


@pytest.mark.parametrize("var", [False, True])
@pytest.mark.parametrize("case", CASES_128, ids=[c[0] for c in CASES_128])
def test_production_size_kernels(case, var, oracle):
    """128^3-class grids: level 0 has 272 smoother CTAs and several interpolation / residual tiles per direction, i.e.
    the kernels and the routing of the 256^3 benchmark (k_smooth_iso, k_residual_iso, k_interp_tile, k_restrict)"""
    from incflo_b200.nodal_projector import A_COR, A_RES, A_RESCOR, OP_INTERP, OP_RESIDUAL, OP_RESTRICT, OP_SMOOTH, SMOOTH_ZERO_START
    mg, proj, rng = _setup(case, var, oracle)
    for lev in range(2):
        phi = _masked_random(mg, lev, rng)
        rhs = _masked_random(mg, lev, rng)
        proj.level_set(lev, A_COR, phi); proj.level_set(lev, A_RES, rhs)
        proj.level_op(lev, OP_SMOOTH, 2)
        _close(proj.level_get(lev, A_COR), mg.smooth(lev, phi.copy(), rhs, 2), 1e-11)
        proj.level_set(lev, A_COR, rng.standard_normal(mg.node_shape(lev)))
        proj.level_op(lev, OP_SMOOTH, 2 | SMOOTH_ZERO_START)
        _close(proj.level_get(lev, A_COR), mg.smooth(lev, np.zeros_like(phi), rhs, 2), 1e-11)
        proj.level_set(lev, A_COR, phi)
        proj.level_op(lev, OP_RESIDUAL)
        r = proj.level_get(lev, A_RESCOR)
        _close(r, mg.residual(lev, phi, rhs))
        proj.level_op(lev, OP_RESTRICT)
        _close(proj.level_get(lev + 1, A_RES), mg.restrict(lev, r))
        crse = _masked_random(mg, lev + 1, rng)
        proj.level_set(lev + 1, A_COR, crse)
        proj.level_op(lev, OP_INTERP)
        _close(proj.level_get(lev, A_COR), mg.interp_add(lev, phi.copy(), crse))
