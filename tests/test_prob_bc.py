"""IncfloVelFill (src/prob/prob_bc.H:8-351): the host restatement against hand-computed values (CPU), and the
device kernel behind b200np_set_inflow_profile against the oracle fed with the host restatement (GPU)."""
import numpy as np
import pytest

from helpers import TILE, oracle_params, rel_l2

INF, NEU, DIR, PER = 3, 1, 2, 0


def test_host_restatement_matches_reference_formulas():
    from incflo_b200 import prob_bc
    n = (8, 4, 6)
    bcv = np.zeros((6, 3)); bcv[0] = (1.0, 0.25, -0.5)          # x-lo: (normal, tangential y, tangential z)
    # channel_cylinder-like deck: probtype 31, parabolic profile 6 y (1 - y) on x-lo (prob_bc.H:62-66)
    a = prob_bc.incflo_vel_fill(n, 2, (INF, NEU, PER), (DIR, NEU, PER), 31, bcv)
    for j in range(4):
        y = (j + 0.5) / 4
        assert np.allclose(a[0, 2:-2, 2 + j, 1], 6 * y * (1 - y))
    assert np.allclose(a[1, 2:-2, 2:-2, 1], 0.25) and np.allclose(a[2, 2:-2, 2:-2, 1], -0.5)
    assert np.count_nonzero(a[:, :, :, 2:]) == 0 and np.count_nonzero(a[:, :, :, 0]) == 0   # only the first ghost layer of x-lo
    # probtype 42: normal velocity = time on both x faces (:57-60, :129-132)
    bcv[3] = (9.0, 0.0, 0.0)
    a = prob_bc.incflo_vel_fill(n, 1, (INF, NEU, NEU), (INF, NEU, NEU), 42, bcv, time=0.375)
    assert np.allclose(a[0, 1:-1, 1:-1, 0], 0.375) and np.allclose(a[0, 1:-1, 1:-1, -1], 0.375)
    # probtype 16 on y-hi (:241-246) and the order of the face blocks at a corner ghost cell (x-lo then y-hi)
    bcv = np.zeros((6, 3)); bcv[0] = (2.0, 3.0, 4.0); bcv[4] = (5.0, 6.0, 7.0)
    a = prob_bc.incflo_vel_fill(n, 1, (INF, NEU, NEU), (NEU, INF, NEU), 16, bcv)
    x = (np.arange(8) + 0.5) / 8
    assert np.allclose(a[1, 3, -1, 1:-1], 16 * (x ** 4 - 2 * x ** 3 + x ** 2))
    assert np.allclose(a[0, 3, -1, 1:-1], 5.0) and np.allclose(a[2, 3, -1, 1:-1], 7.0)
    xg = (-1 + 0.5) / 8                                      # corner (i = -1, j = ny): the y-hi block wrote last
    assert np.isclose(a[1, 3, -1, 0], 16 * (xg ** 4 - 2 * xg ** 3 + xg ** 2)) and np.isclose(a[0, 3, -1, 0], 5.0)
    assert np.isclose(a[0, 3, 2, 0], 2.0) and np.isclose(a[1, 3, 2, 0], 3.0)


@pytest.mark.gpu
@pytest.mark.parametrize("probtype", [31, 311, 0])
def test_device_incflo_vel_fill_matches_oracle_with_host_fill(probtype, oracle):
    """channel: mass inflow x-lo, outflow x-hi, walls y, periodic z; the library fills the inflow ghost layer itself"""
    import torch
    from incflo_b200 import nodal_projector as npj, prob_bc
    n, ng = (48, 16, 16), 2
    dx = (1 / 48,) * 3
    bclo, bchi = (INF, NEU, PER), (DIR, NEU, PER)
    rng = np.random.default_rng(4)
    vel = np.zeros((3, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng))
    vel[:, ng:-ng, ng:-ng, ng:-ng] = 0.1 * rng.standard_normal((3, n[2], n[1], n[0])) + np.array([1.0, 0, 0])[:, None, None, None]
    gp = np.zeros((3, n[2], n[1], n[0])); p = np.zeros((n[2] + 1, n[1] + 1, n[0] + 1))
    bcv = np.zeros((6, 3)); bcv[0] = (1.0, 0.05, -0.02)
    fill = prob_bc.incflo_vel_fill(n, ng, bclo, bchi, probtype, bcv)
    ov, ogp, op_ = vel.copy(), gp.copy(), p.copy()
    status, ost = oracle.apply_nodal_projection(oracle_params(n, dx, bclo, bchi), ov, ng, ogp, op_, inflow_vel=fill,
                                                scaling_factor=0.01)
    assert status == 0
    ip = npj.IncfloProjection(n, dx, bclo, bchi, opts=npj.nodal_proj_opts(tile=TILE))
    ip.set_inflow_profile(probtype, bcv)
    dv, dg, dp = torch.from_numpy(vel).cuda(), torch.from_numpy(gp).cuda(), torch.from_numpy(p).cuda()
    st = ip.apply_nodal_projection(dv, ng, dg, dp, scaling_factor=0.01)
    assert st.status == 0 and abs(st.iters - ost.iters) <= 1
    inner = (slice(None), slice(ng, -ng), slice(ng, -ng), slice(ng, -ng))
    assert rel_l2(dv.cpu().numpy()[inner], ov[inner]) < 1e-9
    assert rel_l2(dg.cpu().numpy(), ogp) < 1e-9 and rel_l2(dp.cpu().numpy(), op_) < 1e-9
    ip.close()


def test_direction_dependent_fill_and_solvability_restatement():
    """benchmark.inout-like deck (test_no_eb_3d/benchmark.inout: probtype 43, dd on both x faces, slip walls): the
    profile 6y(1-y) - 1 changes sign, so each x face has an inflow part (boundary value) and an outflow part (copy of
    the first interior cell, prob_bc.H:106-109 / :160-163); enforceInOutSolvability then balances the fluxes"""
    from incflo_b200 import prob_bc
    n, ng = (8, 8, 4), 2
    dx = (0.25, 0.125, 0.25)
    rng = np.random.default_rng(1)
    vel = rng.standard_normal((3, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng))
    faces = [prob_bc.FACE_DIRECTION_DEPENDENT, 0, 0, prob_bc.FACE_DIRECTION_DEPENDENT, 0, 0]
    bcv = np.zeros((6, 3)); bcv[0] = (0.0, 0.3, -0.2); bcv[3] = (0.0, 0.1, 0.4)
    a = prob_bc.incflo_vel_fill(n, ng, (INF, NEU, NEU), (INF, NEU, NEU), 43, bcv, face_type=faces, vel=vel)
    for j in range(n[1]):
        y = (j + 0.5) / n[1]
        prof = 6 * y * (1 - y) - 1.0
        lo, hi = a[:, ng:-ng, ng + j, ng - 1], a[:, ng:-ng, ng + j, ng + n[0]]
        if prof >= 0:     # x-lo: into the domain -> boundary value (normal: profile, tangential: bcv)
            assert np.allclose(lo[0], prof) and np.allclose(lo[1], 0.3) and np.allclose(lo[2], -0.2)
        else:             # out of the domain -> first interior cell
            assert np.array_equal(lo, vel[:, ng:-ng, ng + j, ng])
        if prof <= 0:     # x-hi: a negative normal velocity points into the domain
            assert np.allclose(hi[0], prof) and np.allclose(hi[1], 0.1) and np.allclose(hi[2], 0.4)
        else:
            assert np.array_equal(hi, vel[:, ng:-ng, ng + j, ng + n[0] - 1])
    assert np.count_nonzero(a[:, ng:-ng, ng:-ng, ng:-ng]) == 0
    # solvability: net flux through the dd faces vanishes afterwards, inflow cells untouched
    b = a.copy()
    fin, fout = prob_bc.enforce_inout_solvability(b, n, ng, dx, faces)
    assert fin > 0 and fout > 0
    ds = dx[1] * dx[2]
    lo, hi = b[0, ng:-ng, ng:-ng, ng - 1], b[0, ng:-ng, ng:-ng, ng + n[0]]
    assert abs((lo.sum() - hi.sum()) * ds) < 1e-12 * fin
    lo0 = a[0, ng:-ng, ng:-ng, ng - 1]
    assert np.array_equal(lo[lo0 >= 0], lo0[lo0 >= 0]) and np.allclose(lo[lo0 < 0], lo0[lo0 < 0] * fin / fout)
    # only inflow -> AMReX-Hydro aborts
    c = np.zeros_like(a); c[0, ng:-ng, ng:-ng, ng - 1] = 1.0
    with pytest.raises(RuntimeError):
        prob_bc.enforce_inout_solvability(c, n, ng, dx, faces)


def test_nodal_bc_mask_restatement():
    """make_nodalBC_mask / prob_set_BC_MF, probtype 1101: x faces mixed, split along y at ny/2"""
    from incflo_b200 import prob_bc
    n = (8, 6, 4)
    faces = [prob_bc.FACE_MIXED, 0, 0, prob_bc.FACE_MIXED, 0, 0]
    m = prob_bc.make_nodalBC_mask(n, faces, 1, n[1] // 2)
    assert m.shape == (5, 7, 9) and m.dtype == np.int32
    assert np.all(m[:, :4, 0] == 0) and np.all(m[:, 4:, 0] == 1)       # x-lo: j <= 3 outflow (prob_bc.cpp:40-44)
    assert np.all(m[:, 4:, -1] == 0) and np.all(m[:, :4, -1] == 1)     # x-hi: j > 3 outflow (:71-75)
    assert np.all(m[:, :, 1:-1] == 1)


@pytest.mark.gpu
@pytest.mark.parametrize("caller_fills", [False, True], ids=["library_fill", "caller_inflow_vel"])
def test_direction_dependent_faces_and_inout_solvability(caller_fills, oracle):
    """test_no_eb_3d/benchmark.inout: dd on x-lo / x-hi with probtype 43, slip walls elsewhere -- a fully Neumann
    (singular) solve that only has a solution after enforceInOutSolvability"""
    import torch
    from incflo_b200 import nodal_projector as npj, prob_bc
    n, ng = (64, 32, 32), 2
    dx = (2.0 / 64, 1.0 / 32, 1.0 / 32)
    bclo, bchi = (INF, NEU, NEU), (INF, NEU, NEU)
    faces = [prob_bc.FACE_DIRECTION_DEPENDENT, 0, 0, prob_bc.FACE_DIRECTION_DEPENDENT, 0, 0]
    rng = np.random.default_rng(8)
    vel = np.zeros((3, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng))
    y = (np.arange(n[1]) + 0.5) / n[1]
    vel[0, ng:-ng, ng:-ng, ng:-ng] = (6 * y * (1 - y) - 1.0)[None, :, None]
    vel[:, ng:-ng, ng:-ng, ng:-ng] += 0.05 * rng.standard_normal((3, n[2], n[1], n[0]))
    gp = 0.02 * rng.standard_normal((3, n[2], n[1], n[0])); p = np.zeros((n[2] + 1, n[1] + 1, n[0] + 1))
    bcv = np.zeros((6, 3))
    dt = 0.01
    # checker: numpy restatement of the ghost fill (on the velocity after u += dt gp / rho, :39-62) + solvability,
    # handed to the oracle as its inflow array
    pre = vel.copy(); pre[:, ng:-ng, ng:-ng, ng:-ng] += dt * gp
    fill = prob_bc.incflo_vel_fill(n, ng, bclo, bchi, 43, bcv, face_type=faces, vel=pre)
    unbalanced = fill.copy()
    fin, fout = prob_bc.enforce_inout_solvability(fill, n, ng, dx, faces)
    ov, ogp, op_ = vel.copy(), gp.copy(), p.copy()
    status, ost = oracle.apply_nodal_projection(oracle_params(n, dx, bclo, bchi), ov, ng, ogp, op_, inflow_vel=fill, scaling_factor=dt)
    assert status == 0
    ip = npj.IncfloProjection(n, dx, bclo, bchi, opts=npj.nodal_proj_opts(tile=TILE))
    ip.set_face_types(faces)
    dv, dg, dp = torch.from_numpy(vel).cuda(), torch.from_numpy(gp).cuda(), torch.from_numpy(p).cuda()
    if caller_fills:     # the caller ran IncfloVelFill itself; the library still balances the fluxes
        st = ip.apply_nodal_projection(dv, ng, dg, dp, inflow_vel=torch.from_numpy(unbalanced).cuda(), scaling_factor=dt)
    else:
        ip.set_inflow_profile(43, bcv)
        st = ip.apply_nodal_projection(dv, ng, dg, dp, scaling_factor=dt)
    gin, gout = ip.inout_flux()
    assert abs(gin - fin) < 1e-12 * fin and abs(gout - fout) < 1e-12 * fout
    assert st.status == 0 and abs(st.iters - ost.iters) <= 1
    inner = (slice(None), slice(ng, -ng), slice(ng, -ng), slice(ng, -ng))
    assert rel_l2(dv.cpu().numpy()[inner], ov[inner]) < 1e-9
    assert rel_l2(dg.cpu().numpy(), ogp) < 1e-9
    a, b = dp.cpu().numpy(), op_
    assert rel_l2(a - a.mean(), b - b.mean()) < 1e-9
    # the ghost layer the library left behind is the balanced fill
    got = dv.cpu().numpy()
    assert rel_l2(got[:, ng:-ng, ng:-ng, ng - 1], fill[:, ng:-ng, ng:-ng, ng - 1]) < 1e-13
    assert rel_l2(got[:, ng:-ng, ng:-ng, ng + n[0]], fill[:, ng:-ng, ng:-ng, ng + n[0]]) < 1e-13
    ip.close()


@pytest.mark.gpu
def test_inflow_without_outflow_is_an_error():
    import torch
    from incflo_b200 import nodal_projector as npj, prob_bc
    n, ng = (16, 16, 16), 1
    ip = npj.IncfloProjection(n, (1 / 16,) * 3, (INF, NEU, NEU), (INF, NEU, NEU))
    ip.set_face_types([prob_bc.FACE_DIRECTION_DEPENDENT, 0, 0, prob_bc.FACE_DIRECTION_DEPENDENT, 0, 0])
    bcv = np.zeros((6, 3)); bcv[0] = (1.0, 0, 0); bcv[3] = (-1.0, 0, 0)     # both faces blow into the domain
    ip.set_inflow_profile(0, bcv)
    vel = torch.zeros((3, 18, 18, 18), dtype=torch.float64, device="cuda")
    gp = torch.zeros((3, 16, 16, 16), dtype=torch.float64, device="cuda"); p = torch.zeros((17, 17, 17), dtype=torch.float64, device="cuda")
    with pytest.raises(npj.ProjectionError) as e:
        ip.apply_nodal_projection(vel, ng, gp, p)
    assert e.value.status == 8    # B200NP_ERR_INOUT_FLUX
    ip.close()


@pytest.mark.gpu
@pytest.mark.parametrize("probtype,split", [(1101, 1), (1102, 2)])
def test_mixed_faces_overset_mask_and_fill(probtype, split, oracle):
    """incflo BC::mixed (probtypes 1101 / 1102): LinOpBCType::inflow + the overset mask of make_nodalBC_mask; the ghost
    velocity comes from the special-case blocks of IncfloVelFill (prob_bc.H:86-92, :140-146, :243-251)"""
    import torch
    from incflo_b200 import nodal_projector as npj, prob_bc
    n, ng = (32, 24, 16), 1
    dx = (1 / 32,) * 3
    if probtype == 1101:
        bclo, bchi = (INF, NEU, NEU), (INF, NEU, NEU)
        faces = [prob_bc.FACE_MIXED, 0, 0, prob_bc.FACE_MIXED, 0, 0]
        mixed = dict(mixed_lo=(1, 0, 0), mixed_hi=(1, 0, 0))
    else:
        bclo, bchi = (NEU, NEU, NEU), (NEU, INF, NEU)
        faces = [0, 0, 0, 0, prob_bc.FACE_MIXED, 0]
        mixed = dict(mixed_lo=(0, 0, 0), mixed_hi=(0, 1, 0))
    half = n[split] // 2
    rng = np.random.default_rng(12)
    vel = np.zeros((3, n[2] + 2, n[1] + 2, n[0] + 2))
    vel[:, 1:-1, 1:-1, 1:-1] = 0.1 * rng.standard_normal((3, n[2], n[1], n[0]))
    gp = np.zeros((3, n[2], n[1], n[0])); p = np.zeros((n[2] + 1, n[1] + 1, n[0] + 1))
    bcv = np.zeros((6, 3)); bcv[0] = (1.0, 0.1, 0.0); bcv[3] = (-0.8, 0.0, 0.1); bcv[4] = (0.0, -0.7, 0.2)
    fill = prob_bc.incflo_vel_fill(n, ng, bclo, bchi, probtype, bcv, face_type=faces, vel=vel)
    assert np.count_nonzero(fill) > 0
    ov, ogp, op_ = vel.copy(), gp.copy(), p.copy()
    prm = oracle_params(n, dx, bclo, bchi, mix_dir=split, mix_half=half, **mixed)
    status, ost = oracle.apply_nodal_projection(prm, ov, ng, ogp, op_, inflow_vel=fill, scaling_factor=0.01)
    assert status == 0
    ip = npj.IncfloProjection(n, dx, bclo, bchi, opts=npj.nodal_proj_opts(tile=TILE))
    ip.set_face_types(faces, split, half)
    mask = prob_bc.make_nodalBC_mask(n, faces, split, half)
    assert ip.check_overset_mask(mask) and ip.check_overset_mask(torch.from_numpy(mask).cuda())
    wrong = mask.copy(); wrong[3, 3, 3] = 0
    assert not ip.check_overset_mask(wrong)
    ip.set_inflow_profile(probtype, bcv)
    dv, dg, dp = torch.from_numpy(vel).cuda(), torch.from_numpy(gp).cuda(), torch.from_numpy(p).cuda()
    st = ip.apply_nodal_projection(dv, ng, dg, dp, scaling_factor=0.01)
    assert st.status == 0 and abs(st.iters - ost.iters) <= 1
    inner = (slice(None), slice(1, -1), slice(1, -1), slice(1, -1))
    assert rel_l2(dv.cpu().numpy()[inner], ov[inner]) < 1e-9
    assert rel_l2(dg.cpu().numpy(), ogp) < 1e-9 and rel_l2(dp.cpu().numpy(), op_) < 1e-9
    # pressure is pinned to zero on the outflow half of the mixed faces
    assert np.all(dp.cpu().numpy()[mask == 0] == 0.0)
    ip.close()
