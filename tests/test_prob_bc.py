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
