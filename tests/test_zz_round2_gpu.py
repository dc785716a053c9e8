"""GPU tests of code written AFTER the last GPU visit of round 1 (no B200 minutes were left to run them).
They are xfail(strict=False): a pass is reported as XPASS, a wrong result does not turn the suite red.  The file
sorts last on purpose, so that nothing here can disturb the verified GPU tests that run before it.  Round 2:
run them, fix what fails, and move them next to their verified siblings."""
import os

import numpy as np
import pytest

from helpers import TILE, oracle_params, rel_l2
from test_composite_oracle import GOLD_DIR

INF, NEU, DIR, PER = 3, 1, 2, 0


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="non-singular composite problems (an outflow face): first GPU run is in round 2")
@pytest.mark.parametrize("host", [False, True], ids=["device_ptrs", "host_ptrs"])
@pytest.mark.parametrize("path", GOLD_DIR, ids=[os.path.basename(p)[:-4] for p in GOLD_DIR])
def test_composite_cuda_reproduces_nonsingular_golden(path, host):
    from test_gpu_composite import _golden_case
    _golden_case(path, host)


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="k_incflo_vel_fill was written after the last GPU visit of round 1")
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
