"""End-to-end parity of the CUDA projection (through the C ABI) against the CPU oracle.

north_star tolerances, written out:
  * residual criterion  ||rhs - L phi||_inf <= max(atol, rtol * max(||rhs||, ||res0||))   (A.9)
  * pressure (phi, mean removed when the problem is singular) and projected velocity within
    1e-9 relative L2 of the oracle at rtol = 1e-11
  * V-cycle count within 20 % of the oracle (AMReX CPU algorithm: lexicographic Gauss-Seidel)
"""
import numpy as np
import pytest

from helpers import BC_CASES, TILE, oracle_params, rel_l2, remove_mean

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-11, 1e-14
PARITY = 1e-9


def _np(t):
    return t.detach().cpu().numpy().copy()


def _run_gpu(cfg, host, **optkw):
    import torch
    from incflo_b200 import nodal_projector as npj
    vel = cfg["vel"].clone()
    sigma = cfg["sigma"]
    if host:
        vel_in = _np(vel); sig_in = None if sigma is None else _np(sigma)
    else:
        vel_in = vel.cuda(); sig_in = None if sigma is None else sigma.cuda()
    proj = npj.NodalProjector(vel_in, sig_in, cfg["const_sigma"],
                              dict(n_cell=cfg["n"], dx=cfg["dx"], is_periodic=[b == 0 for b in cfg["bclo"]]),
                              ng=1, opts=npj.nodal_proj_opts(tile=TILE, **optkw))
    proj.setDomainBC(cfg["bclo"], cfg["bchi"])
    st = proj.project(RTOL, ATOL)
    phi, g = proj.getPhi(), proj.getGradPhi()
    if not host:
        torch.cuda.synchronize()
        vel_in, phi, g = _np(vel_in), _np(phi), _np(g)
    proj.close()
    return vel_in, phi, g, st


def _run_oracle(cfg, oracle, mirror=True, **kw):
    vel = _np(cfg["vel"])
    sigma = None if cfg["sigma"] is None else _np(cfg["sigma"])
    if mirror:
        p = oracle_params(cfg["n"], cfg["dx"], cfg["bclo"], cfg["bchi"], **kw)
    else:
        p = oracle.make_params(cfg["n"], cfg["dx"], cfg["bclo"], cfg["bchi"], smoother=oracle.SM_LEX, **kw)
    r = oracle.project(p, vel, 1, sigma, cfg["const_sigma"], RTOL, ATOL, want_rhs=True)
    return vel, r


@pytest.mark.parametrize("host", [False, True], ids=["device_ptrs", "host_ptrs"])
@pytest.mark.parametrize("config,N", [("tgv", 32), ("tgv", 64), ("rt", 32), ("rt", 64), ("dsl", 64)])
def test_project_parity(config, N, host, oracle):
    from incflo_b200 import problems
    cfg = problems.make(config, N, ng=1, device="cpu")
    gvel, gphi, ggrad, st = _run_gpu(cfg, host)
    # (1) converged by the reference criterion
    assert st.status == 0
    assert st.resnorm <= max(ATOL, RTOL * max(st.rhsnorm, st.resnorm0))
    # (2) same algorithm on the CPU (oracle configured with the GPU's tile ordering)
    ovel, r = _run_oracle(cfg, oracle, mirror=True)
    ost = r["stats"]
    assert ost.status == 0
    assert st.iters == ost.iters, (st.iters, ost.iters)
    assert abs(st.rhsnorm - ost.rhsnorm) <= 1e-12 * ost.rhsnorm
    assert rel_l2(remove_mean(gphi), remove_mean(r["phi"])) < PARITY
    assert rel_l2(ggrad, r["gphi"]) < PARITY
    ng = 1
    inner = (slice(None), slice(ng, -ng), slice(ng, -ng), slice(ng, -ng))
    assert rel_l2(gvel[inner], ovel[inner]) < PARITY
    # ghost cells of vel are inputs and must be untouched
    mask = np.ones(gvel.shape, bool); mask[inner] = False
    assert np.array_equal(gvel[mask], _np(cfg["vel"])[mask])
    # (3) V-cycle count within 20 % of the reference CPU algorithm (lexicographic GS, one box)
    _, rl = _run_oracle(cfg, oracle, mirror=False)
    assert rl["stats"].status == 0
    assert abs(st.iters - rl["stats"].iters) <= 0.2 * rl["stats"].iters + 1e-9, (st.iters, rl["stats"].iters)
    assert rel_l2(remove_mean(gphi), remove_mean(rl["phi"])) < 1e-7  # different smoother, same discrete solution


@pytest.mark.parametrize("top", ["direct", "correction"])
@pytest.mark.parametrize("case", BC_CASES, ids=[c[0] for c in BC_CASES])
@pytest.mark.parametrize("var", [False, True])
def test_project_bc_cases(case, var, top, oracle, monkeypatch):
    """random velocity + every BC combination (walls, inflow with non-zero ghost velocity,
    Dirichlet outflow, anisotropic dx), against the mirrored oracle.  `top`: the finest level of the V-cycle relaxes
    (sol, rhs) directly (default) or MLMG's (cor, res) followed by sol += cor -- the same iterates up to rounding."""
    import torch
    monkeypatch.setenv("B200NP_TOP_DIRECT", "1" if top == "direct" else "0")
    name, n, dx, bclo, bchi = case
    rng = np.random.default_rng(7)
    vel = rng.standard_normal((3, n[2] + 2, n[1] + 2, n[0] + 2))
    # smooth it a little so the multigrid sees a realistic spectrum
    for ax in (1, 2, 3):
        vel = 0.5 * vel + 0.25 * (np.roll(vel, 1, ax) + np.roll(vel, -1, ax))
    # ghost cells: 0 at walls, inflow value at inflow faces, garbage at periodic faces (must be ignored)
    for d, ax in ((0, 3), (1, 2), (2, 1)):
        for side, bc in ((0, bclo[d]), (1, bchi[d])):
            sl = [slice(None)] * 4
            sl[ax] = 0 if side == 0 else -1
            if bc == 0:
                vel[tuple(sl)] = 1e30
            elif bc == 3:
                vel[tuple(sl)] = 0.3
            else:
                vel[tuple(sl)] = 0.0
    sigma = np.ascontiguousarray(rng.uniform(0.5, 2.0, size=(n[2], n[1], n[0]))) if var else None
    cfg = dict(vel=torch.from_numpy(vel.copy()), sigma=None if sigma is None else torch.from_numpy(sigma),
               const_sigma=0.37, bclo=bclo, bchi=bchi, n=n, dx=dx)
    gvel, gphi, ggrad, st = _run_gpu(cfg, host=False)
    assert st.status == 0
    ovel, r = _run_oracle(cfg, oracle, mirror=True)
    assert r["stats"].status == 0
    assert st.iters == r["stats"].iters
    singular = all(b != 2 for b in bclo + bchi)
    a, b = (remove_mean(gphi), remove_mean(r["phi"])) if singular else (gphi, r["phi"])
    assert rel_l2(a, b) < PARITY
    assert rel_l2(ggrad, r["gphi"]) < PARITY
    inner = (slice(None), slice(1, -1), slice(1, -1), slice(1, -1))
    assert rel_l2(gvel[inner], ovel[inner]) < PARITY


@pytest.mark.parametrize("use_graph", [0, 1])
def test_graph_and_repeat(use_graph, oracle):
    """the cached handle / CUDA graph give identical results on repeated calls and when the
    constant sigma changes between calls (time step change)"""
    from incflo_b200 import nodal_projector as npj, problems
    cfg = problems.make("tgv", 32, ng=1, device="cpu")
    outs = []
    for cs in (cfg["const_sigma"], cfg["const_sigma"], 2.5 * cfg["const_sigma"]):
        vel = _np(cfg["vel"])
        proj = npj.NodalProjector(vel, None, cs, dict(n_cell=cfg["n"], dx=cfg["dx"], is_periodic=(1, 1, 1)), ng=1,
                                  opts=npj.nodal_proj_opts(tile=TILE, use_graph=use_graph))
        proj.project(RTOL, ATOL)
        outs.append((vel, proj.getPhi().copy()))
        # second projection on the same handle with another sigma
        vel2 = _np(cfg["vel"])
        proj.vel = vel2; proj.const_sigma = 2.5 * cs
        proj.project(RTOL, ATOL)
        outs.append((vel2, proj.getPhi().copy()))
        proj.close()
    assert np.array_equal(outs[0][0], outs[2][0]) and np.array_equal(outs[0][1], outs[2][1])  # deterministic
    assert np.array_equal(outs[1][0], outs[4][0])  # handle reuse == fresh handle with that sigma
    # phi scales like 1/sigma, velocity is independent of a constant sigma
    assert rel_l2(outs[1][1] * 2.5, outs[0][1]) < 1e-9
    assert rel_l2(outs[1][0], outs[0][0]) < 1e-9


@pytest.mark.parametrize("incremental,small_dt,var", [(False, False, True), (True, False, True), (False, True, False),
                                                      (False, False, False)])
def test_apply_nodal_projection(incremental, small_dt, var, oracle):
    """incflo::ApplyNodalProjection semantics (:29-267): gp pre-add, sigma = s/rho, u -= u_old,
    setBndry(0), copy-out / accumulate of gp and p_nd."""
    from incflo_b200 import nodal_projector as npj, problems
    N, ng = 32, 3
    n = (N, N, N)
    bclo, bchi = (0, 0, 1), (0, 0, 1)
    rng = np.random.default_rng(3)
    vel = _np(problems.rayleigh_taylor_velocity(n, ng, "cpu", "b"))
    vel[:, :ng] = 7.0  # junk in ghost cells: must be zeroed by setBndry
    velo = 0.9 * vel + 0.01 * rng.standard_normal(vel.shape)
    rho = _np(problems.rayleigh_taylor_density(n, ng, "cpu")) if var else None
    gp = 0.1 * rng.standard_normal((3, N, N, N))
    p = rng.standard_normal((N + 1, N + 1, N + 1))
    dt = 0.01
    args = dict(ro_0=1.3, scaling_factor=dt, incremental=incremental, proj_for_small_dt=small_dt)
    # oracle
    ov, ogp, op_ = vel.copy(), gp.copy(), p.copy()
    prm = oracle_params(n, (1 / N,) * 3, bclo, bchi)
    status, ost = oracle.apply_nodal_projection(prm, ov, ng, ogp, op_, density=rho, ngd=ng, velocity_o=velo,
                                                rtol=RTOL, atol=ATOL, **args)
    assert status == 0
    # CUDA through the C ABI (host pointers => staging inside the call)
    gv, ggp, gpn = vel.copy(), gp.copy(), p.copy()
    ip = npj.IncfloProjection(n, (1 / N,) * 3, bclo, bchi, opts=npj.nodal_proj_opts(tile=TILE))
    st = ip.apply_nodal_projection(gv, ng, ggp, gpn, density=rho, ngd=ng, velocity_o=velo, mg_rtol=RTOL, mg_atol=ATOL,
                                   **args)
    ip.close()
    assert st.iters == ost.iters
    inner = (slice(None), slice(ng, -ng), slice(ng, -ng), slice(ng, -ng))
    assert rel_l2(gv[inner], ov[inner]) < PARITY
    assert rel_l2(ggp, ogp) < PARITY
    if incremental:  # p_nd += phi
        assert rel_l2(remove_mean(gpn - p), remove_mean(op_ - p)) < PARITY
    else:            # p_nd = phi
        assert rel_l2(remove_mean(gpn), remove_mean(op_)) < PARITY
    mask = np.ones(gv.shape, bool); mask[inner] = False
    assert np.all(gv[mask] == 0.0)
    assert st.h2d_bytes > 0 and st.d2h_bytes > 0


def test_properties_at_baseline_size():
    """BASELINE configs[1] at its FULL size (rayleigh_taylor 256^3, variable sigma = dt / rho; the CPU oracle would need ~20 s and 16 threads,
    bench.py runs that comparison at 128^3) through properties that need no oracle: convergence to mg_rtol, the projected field is
    nearly divergence-free in the nodal sense (approximate projection: a second projection finds a right-hand side O(h) of the first), linearity of the projection in u, phi ~ 1 / sigma at fixed u (sigma -> 4 sigma: same u, phi / 4), and the V-cycle count of
    the benchmark line."""
    import torch
    from incflo_b200 import nodal_projector as npj, problems
    N, ng = 256, 1
    n = (N, N, N)
    dev = torch.device("cuda:0")
    proj = npj.IncfloProjection(n, (1.0 / N,) * 3, (0, 0, 1), (0, 0, 1))
    u0 = torch.zeros((3, N + 2 * ng, N + 2 * ng, N + 2 * ng), dtype=torch.float64, device=dev)
    u0[:, ng:-ng, ng:-ng, ng:-ng] = problems.rayleigh_taylor_velocity(n, 0, dev, "b")
    rho = torch.ones((N + 2 * ng,) * 3, dtype=torch.float64, device=dev)
    rho[ng:-ng, ng:-ng, ng:-ng] = problems.rayleigh_taylor_density(n, 0, dev)
    dt = 0.45 / N
    inner = (slice(None), slice(ng, ng + N), slice(ng, ng + N), slice(ng, ng + N))

    def rel(a, b):
        return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))

    def run(vel, scale_u=1.0, sigma_scale=1.0, rtol=1e-11):
        v = (scale_u * vel).clone()
        gp = torch.zeros((3, N, N, N), dtype=torch.float64, device=dev)
        p = torch.zeros((N + 1,) * 3, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()   # the handle runs on its own stream (b200np_set_stream not used here): torch's fills must have landed
        st = proj.apply_nodal_projection(v, ng, gp, p, density=rho, ngd=ng, scaling_factor=dt * sigma_scale, mg_rtol=rtol, mg_atol=1e-14)
        assert st.status == 0 and st.resnorm <= max(1e-14, rtol * max(st.rhsnorm, st.resnorm0))
        return v[inner].clone(), p, gp, st.iters, st.rhsnorm

    u1, p1, g1, it1, rhs1 = run(u0)
    assert 6 <= it1 <= 9                                      # the bench line: 7 V-cycles
    # the projection is APPROXIMATE (the update uses the cell-averaged gradient, the operator the full trilinear one): the nodal divergence
    # of the projected field is not zero but O(h) of the original one -- 1.9 % at 32^3 in the CPU oracle, 8 x less here
    v2 = torch.zeros_like(u0); v2[inner] = u1
    u2, p2, g2, it2, rhs2 = run(v2, rtol=1e-3)
    assert rhs2 <= 5e-3 * rhs1
    # linearity in u (p = phi / dt in the non-incremental form, gp = grad phi: both linear)
    u3, p3, g3, _, _ = run(u0, scale_u=2.5)
    assert rel(u3, 2.5 * u1) < 1e-9 and rel(p3, 2.5 * p1) < 1e-9 and rel(g3, 2.5 * g1) < 1e-9
    # sigma -> 4 sigma at fixed u (scaling_factor = 4 dt, gp_old = 0): the same projected velocity, p_nd = phi and gp = grad phi shrink by 4
    u4, p4, g4, _, _ = run(u0, sigma_scale=4.0)
    assert rel(u4, u1) < 1e-9 and rel(4.0 * p4, p1) < 1e-9 and rel(4.0 * g4, g1) < 1e-9
    proj.close()
