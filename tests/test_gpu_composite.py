"""Composite (two AMR level, one fine box at ratio 2; BASELINE configs[3]) projection on the GPU through
b200np_composite_* against (1) the golden composite finite-element solutions and (2) the CPU oracle.
Tolerances (north_star): pressure / gradient / projected velocity within 1e-9 relative L2; iteration
count within 20 % of the oracle's (same algorithm and smoother ordering: identical in practice)."""
import glob
import os

import numpy as np
import pytest

from helpers import TILE, oracle_params, rel_l2
from test_composite_oracle import GOLD, GOLD_DIR, check, fine_per, load, to_full

pytestmark = pytest.mark.gpu
MIRROR = None


def _mirror(oracle):
    return dict(smoother=oracle.SM_BOX, box=TILE, box_order=oracle.SM_PLANE4, box_stale_per_call=0)


def _cuda(a, host):
    import torch
    return a.copy() if host else torch.from_numpy(a.copy()).cuda()


def _np(a):
    return a if isinstance(a, np.ndarray) else a.detach().cpu().numpy()


@pytest.mark.parametrize("host", [False, True], ids=["device_ptrs", "host_ptrs"])
@pytest.mark.parametrize("path", GOLD + GOLD_DIR, ids=[os.path.basename(p)[:-4] for p in GOLD + GOLD_DIR])
def test_composite_cuda_reproduces_golden(path, host):
    """singular (periodic / wall) and non-singular (outflow face, also with the box ON the outflow face) fixtures"""
    from incflo_b200 import nodal_projector as npj
    g = load(path)
    cp = npj.CompositeProjection(g["n0"], g["dx0"], g["bclo"], g["bchi"], g["clo"], g["chi"], opts=npj.nodal_proj_opts(tile=TILE))
    v0, v1 = _cuda(g["vel0_in"], host), _cuda(g["vel1_in"], host)
    s0 = _cuda(g["sigma0"], host) if g["var"] else None
    s1 = _cuda(g["sigma1"], host) if g["var"] else None
    phi0, phi1, g0, g1 = cp.project(v0, g["ng0"], v1, g["ng1"], s0, s1, float(g["sigma0"].flat[0]), rtol=1e-12, atol=0.0)
    assert cp.stats.status == 0 and cp.stats.iters <= 25
    check(g, _np(v0), _np(v1), _np(phi0), _np(phi1), _np(g0), _np(g1))
    cp.close()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_composite_cuda_matches_oracle_iteration(path, oracle):
    """same input, same algorithm: same iteration count, residual history and solution"""
    from incflo_b200 import nodal_projector as npj
    from oracle import composite as oc
    g = load(path)
    p0 = oracle_params(g["n0"], g["dx0"], g["bclo"], g["bchi"])
    ocp = oc.CompositeProjector(p0, g["clo"], g["chi"], smoother_kw=_mirror(oracle))
    ov0, ov1 = g["vel0_in"].copy(), g["vel1_in"].copy()
    r = ocp.project(ov0, g["ng0"], ov1, g["ng1"], g["sigma0"] if g["var"] else None, g["sigma1"] if g["var"] else None,
                    float(g["sigma0"].flat[0]), rtol=1e-11, atol=1e-14)
    cp = npj.CompositeProjection(g["n0"], g["dx0"], g["bclo"], g["bchi"], g["clo"], g["chi"], opts=npj.nodal_proj_opts(tile=TILE))
    v0, v1 = _cuda(g["vel0_in"], False), _cuda(g["vel1_in"], False)
    s0 = _cuda(g["sigma0"], False) if g["var"] else None
    s1 = _cuda(g["sigma1"], False) if g["var"] else None
    phi0, phi1, g0, g1 = cp.project(v0, g["ng0"], v1, g["ng1"], s0, s1, float(g["sigma0"].flat[0]), rtol=1e-11, atol=1e-14)
    st = cp.stats
    assert st.status == 0 and r["status"] == 0
    assert abs(st.iters - r["iters"]) <= max(1, int(0.2 * r["iters"]))
    assert abs(st.rhsnorm - r["rhsnorm"]) <= 1e-10 * r["rhsnorm"]
    c = _np(phi1).mean() - to_full(r["phi1"], fine_per(g)).mean()
    assert rel_l2(_np(phi1) - c, to_full(r["phi1"], fine_per(g))) < 1e-9
    assert rel_l2(_np(phi0) - c, to_full(r["phi0"], g["bclo"])) < 1e-9
    assert rel_l2(_np(g1), r["gphi1"]) < 1e-9 and rel_l2(_np(g0), r["gphi0"]) < 1e-9
    n0 = g["n0"]; a = g["ng0"]
    i0 = (slice(None), slice(a, a + n0[2]), slice(a, a + n0[1]), slice(a, a + n0[0]))
    assert rel_l2(_np(v0)[i0], ov0[i0]) < 1e-9
    cp.close()


def test_composite_apply_nodal_projection_variable_density(oracle):
    """incflo::ApplyNodalProjection with finest_level = 1 (pre-add u += dt gp / rho, sigma = dt / rho per level,
    copy-out gp / p_nd, average_down) against the oracle's restatement; bouss_bubble-like geometry at 32^3 / box 16^3"""
    import torch
    from incflo_b200 import nodal_projector as npj
    from oracle import composite as oc
    N, ng = 32, (3, 3)
    n0 = (N, N, N); dx0 = (1.0 / N,) * 3
    bclo = bchi = (0, 0, 1)
    clo, chi = (8, 8, 8), (23, 23, 23)
    nf = (32, 32, 32)
    rng = np.random.default_rng(12)

    def smooth(shape):
        v = rng.standard_normal(shape)
        for ax in range(v.ndim - 3, v.ndim):
            for _ in range(2):
                v = 0.5 * v + 0.25 * (np.roll(v, 1, ax) + np.roll(v, -1, ax))
        return v
    vel, gp, p, rho = [], [], [], []
    for n in (n0, nf):
        v = np.zeros((3, n[2] + 6, n[1] + 6, n[0] + 6)); v[:, 3:-3, 3:-3, 3:-3] = smooth((3,) + n[::-1])
        vel.append(v)
        gp.append(0.1 * smooth((3,) + n[::-1]))
        p.append(np.zeros((n[2] + 1, n[1] + 1, n[0] + 1)))
        r = np.ones((n[2] + 6, n[1] + 6, n[0] + 6)); r[3:-3, 3:-3, 3:-3] = 1.0 + 0.5 * np.tanh(4 * smooth(n[::-1]))
        rho.append(r)
    dt = 0.45 / N
    ovel = [v.copy() for v in vel]; ogp = [x.copy() for x in gp]; op_ = [x.copy() for x in p]
    ocp = oc.CompositeProjector(oracle_params(n0, dx0, bclo, bchi), clo, chi, smoother_kw=_mirror(oracle))
    r = oc.apply_nodal_projection(ocp, ovel, ng, ogp, op_, density=rho, ngd=ng, scaling_factor=dt)
    assert r["status"] == 0
    cp = npj.CompositeProjection(n0, dx0, bclo, bchi, clo, chi, opts=npj.nodal_proj_opts(tile=TILE))
    dvel = [torch.from_numpy(v).cuda() for v in vel]; dgp = [torch.from_numpy(x).cuda() for x in gp]
    dp = [torch.from_numpy(x).cuda() for x in p]; drho = [torch.from_numpy(x).cuda() for x in rho]
    st = cp.apply_nodal_projection(dvel, ng, dgp, dp, density=drho, ngd=ng, scaling_factor=dt)
    assert st.status == 0 and abs(st.iters - r["iters"]) <= max(1, int(0.2 * r["iters"]))
    c = _np(dp[1]).mean() - op_[1].mean()
    for l, n in enumerate((n0, nf)):
        inner = (slice(None), slice(3, 3 + n[2]), slice(3, 3 + n[1]), slice(3, 3 + n[0]))
        assert rel_l2(_np(dvel[l])[inner], ovel[l][inner]) < 1e-9
        assert rel_l2(_np(dgp[l]), ogp[l]) < 1e-9
        assert rel_l2(_np(dp[l]) - c, op_[l]) < 1e-9
    # host pointers through the same entry point
    hvel = [v.copy() for v in vel]; hgp = [x.copy() for x in gp]; hp = [x.copy() for x in p]
    st = cp.apply_nodal_projection(hvel, ng, hgp, hp, density=rho, ngd=ng, scaling_factor=dt)
    assert st.status == 0 and st.h2d_bytes > 0 and st.d2h_bytes > 0
    for l in range(2):
        assert rel_l2(hgp[l], _np(dgp[l])) < 1e-13 and rel_l2(hp[l], _np(dp[l])) < 1e-13
    cp.close()


@pytest.mark.parametrize("host", [True, False], ids=["host_ptrs", "device_ptrs"])
@pytest.mark.parametrize("max_grid", [16, 12], ids=["grid16", "ragged12"])
def test_composite_multibox_equals_single_box(max_grid, host):
    """b200np_composite_apply_nodal_projection_mf: both AMR levels chopped into boxes of amr.max_grid_size cells (level 1 in fine index
    space, its boxes tiling the fine box) must reproduce the single-box call bit for bit (velocity, gp, p_nd on both levels)"""
    import torch
    from incflo_b200 import nodal_projector as npj
    N, ng = 32, 2
    n0 = (N, N, N); dx0 = (1.0 / N,) * 3
    bclo = bchi = (0, 0, 1)
    clo, chi = (8, 4, 6), (23, 19, 25)
    nf = tuple(2 * (h - l + 1) for l, h in zip(clo, chi))
    org = tuple(2 * l for l in clo)
    rng = np.random.default_rng(21)
    vel, gp, p, rho = [], [], [], []
    for n in (n0, nf):
        v = np.zeros((3, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng)); v[:, ng:-ng, ng:-ng, ng:-ng] = rng.standard_normal((3,) + n[::-1])
        vel.append(v)
        gp.append(0.1 * rng.standard_normal((3,) + n[::-1]))
        p.append(np.zeros((n[2] + 1, n[1] + 1, n[0] + 1)))
        r = np.ones((n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng)); r[ng:-ng, ng:-ng, ng:-ng] = rng.uniform(1.0, 2.0, size=n[::-1])
        rho.append(r)
    dt = 0.3 / N
    cp = npj.CompositeProjection(n0, dx0, bclo, bchi, clo, chi, opts=npj.nodal_proj_opts(tile=TILE))
    svel = [v.copy() for v in vel]; sgp = [x.copy() for x in gp]; sp = [x.copy() for x in p]
    st1 = cp.apply_nodal_projection(svel, (ng, ng), sgp, sp, density=rho, ngd=(ng, ng), scaling_factor=dt)
    assert st1.status == 0
    iters1 = st1.iters
    to = None if host else (lambda a: torch.from_numpy(a).cuda())
    origins = ((0, 0, 0), org)
    mv = [npj.MultiFab.split(vel[l], (n0, nf)[l], max_grid, ng, 3, to=to, origin=origins[l]) for l in range(2)]
    mg = [npj.MultiFab.split(gp[l], (n0, nf)[l], max_grid, 0, 3, to=to, origin=origins[l]) for l in range(2)]
    mp = [npj.MultiFab.split(p[l][None], (n0, nf)[l], max_grid, 0, 1, nodal=True, to=to, origin=origins[l]) for l in range(2)]
    mr = [npj.MultiFab.split(rho[l][None], (n0, nf)[l], max_grid, ng, 1, to=to, origin=origins[l]) for l in range(2)]
    assert len(mv[1].boxes) > 1 and len(mv[0].boxes) > 1
    st2 = cp.apply_nodal_projection_mf(mv, mg, mp, density=mr, scaling_factor=dt)
    assert st2.status == 0 and st2.iters == iters1
    if host:
        assert st2.h2d_bytes > 0 and st2.d2h_bytes > 0
    for l, n in enumerate((n0, nf)):
        inner = (slice(None), slice(ng, ng + n[2]), slice(ng, ng + n[1]), slice(ng, ng + n[0]))
        assert np.array_equal(mv[l].assemble(n, origin=origins[l]), svel[l][inner])
        assert np.array_equal(mg[l].assemble(n, origin=origins[l]), sgp[l])
        assert np.array_equal(mp[l].assemble(n, origin=origins[l])[0], sp[l])
    # a level-1 MultiFab that does not tile the fine box is refused
    bad = npj.MultiFab(mv[1].boxes[:-1], mv[1].arrays[:-1], ng, 3)
    with pytest.raises(npj.ProjectionError) as e:
        cp.apply_nodal_projection_mf([mv[0], bad], mg, mp, density=mr, scaling_factor=dt)
    assert e.value.status == 4    # B200NP_ERR_BAD_ARG
    cp.close()


def test_composite_rejects_unsupported_boxes():
    from incflo_b200 import nodal_projector as npj
    with pytest.raises(npj.ProjectionError) as e:   # touches the domain face: B200NP_ERR_UNSUPPORTED
        npj.CompositeProjection((16, 16, 16), (1 / 16,) * 3, (0, 0, 1), (0, 0, 1), (0, 4, 4), (7, 11, 11))
    assert e.value.status == 7
