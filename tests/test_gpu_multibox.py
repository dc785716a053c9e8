"""Multi-box MultiFabs at the boundary (b200np_*_mf): every reference deck runs with amr.max_grid_size = 16
(test_no_eb_3d/benchmark.rayleigh_taylor:16).  The reference's own decomposition-invariance pattern: the same problem
cut into boxes must give the same answer -- here bit for bit, because the boxes are gathered into the very arrays
the single-box path works on."""
import numpy as np
import pytest

from helpers import TILE

pytestmark = pytest.mark.gpu


def _problem(n, ng, var, seed=5):
    from incflo_b200 import problems
    rng = np.random.default_rng(seed)
    vel = problems.rayleigh_taylor_velocity(n, ng, "cpu", "b").numpy().copy()
    vel += 0.01 * rng.standard_normal(vel.shape)
    rho = problems.rayleigh_taylor_density(n, ng, "cpu").numpy().copy() if var else None
    gp = 0.1 * rng.standard_normal((3, n[2], n[1], n[0]))
    p = rng.standard_normal((n[2] + 1, n[1] + 1, n[0] + 1))
    velo = 0.9 * vel + 0.01 * rng.standard_normal(vel.shape)
    return vel, velo, rho, gp, p


@pytest.mark.parametrize("device_ptrs", [True, False], ids=["device_ptrs", "host_ptrs"])
@pytest.mark.parametrize("incremental,var", [(False, True), (True, True), (False, False)])
def test_apply_nodal_projection_16cube_boxes_equal_single_box(incremental, var, device_ptrs):
    import torch
    from incflo_b200 import nodal_projector as npj
    N, ng, mg = 64, 3, 16
    n = (N, N, N)
    bclo, bchi = (0, 0, 1), (0, 0, 1)
    vel, velo, rho, gp, p = _problem(n, ng, var)
    args = dict(ro_0=1.3, scaling_factor=0.01, incremental=incremental)
    ip = npj.IncfloProjection(n, (1 / N,) * 3, bclo, bchi, opts=npj.nodal_proj_opts(tile=TILE))
    # single box
    sv, sg, sp = vel.copy(), gp.copy(), p.copy()
    st1 = ip.apply_nodal_projection(sv, ng, sg, sp, density=rho, ngd=ng, velocity_o=velo, **args)
    it1, rhs1 = st1.iters, st1.rhsnorm
    # 64 boxes of 16^3, each with its own ghost frame
    to = (lambda a: torch.from_numpy(a).cuda()) if device_ptrs else None
    mv = npj.MultiFab.split(vel, n, mg, ng, 3, to=to)
    mo = npj.MultiFab.split(velo, n, mg, ng, 3, to=to)
    mr = npj.MultiFab.split(rho, n, mg, ng, 1, to=to) if var else None
    mgp = npj.MultiFab.split(gp, n, mg, 0, 3, to=to)
    mp = npj.MultiFab.split(p, n, mg, 0, 1, nodal=True, to=to)
    assert len(mv.boxes) == 64
    st2 = ip.apply_nodal_projection_mf(mv, mgp, mp, density=mr, velocity_o=mo, **args)
    assert st2.iters == it1 and st2.rhsnorm == rhs1          # the rhs is the same to the last bit
    inner = (slice(None), slice(ng, -ng), slice(ng, -ng), slice(ng, -ng))
    assert np.array_equal(mv.assemble(n), sv[inner])
    assert np.array_equal(mgp.assemble(n), sg)
    assert np.array_equal(mp.assemble(n)[0], sp)
    # ghost cells of a box: inside the domain the neighbours' projected values (as after FillBoundary), outside the
    # domain the BC value in the first layer and zero beyond (setBndry(0), :137)
    a = mv.arrays[21]                      # interior box (16..31)^3
    a = a.cpu().numpy() if device_ptrs else a
    lo, hi = mv.boxes[21]
    assert np.array_equal(a, sv[:, lo[2]:hi[2] + 1 + 2 * ng, lo[1]:hi[1] + 1 + 2 * ng, lo[0]:hi[0] + 1 + 2 * ng])
    b = mv.arrays[0]                       # corner box at z = 0 (wall): ghost layers below the wall
    b = b.cpu().numpy() if device_ptrs else b
    assert np.all(b[:, :ng] == 0.0)        # wall ghost value 0, and zero beyond the first layer
    assert np.all(b[:, ng:-ng, ng:-ng, :ng - 1] == 0.0) and np.all(b[:, ng:-ng, :ng - 1, ng:-ng] == 0.0)   # periodic x / y ghosts beyond layer 1
    ip.close()


def test_project_mf_ragged_boxes_and_inflow_ghosts():
    """NodalProjector::project over boxes that do not divide the domain evenly, on a channel with an inflow face: the BC
    ghost layer of vel travels with the boxes that touch the face"""
    import torch
    from incflo_b200 import nodal_projector as npj
    n, ng, mg = (48, 20, 24), 1, 16
    dx = (1 / 48,) * 3
    bclo, bchi = (3, 1, 0), (2, 1, 0)
    rng = np.random.default_rng(9)
    vel = np.zeros((3, n[2] + 2, n[1] + 2, n[0] + 2))
    vel[:, 1:-1, 1:-1, 1:-1] = 0.1 * rng.standard_normal((3, n[2], n[1], n[0])) + np.array([1.0, 0, 0])[:, None, None, None]
    vel[0, :, :, 0] = 1.0 + 0.1 * rng.standard_normal(vel[0, :, :, 0].shape)      # inflow ghost layer
    sigma = rng.uniform(0.5, 2.0, size=(n[2], n[1], n[0]))
    single = npj.NodalProjector(vel.copy(), sigma.copy(), None, dict(n_cell=n, dx=dx, is_periodic=(0, 0, 1)), ng=ng,
                                opts=npj.nodal_proj_opts(tile=TILE))
    single.setDomainBC(bclo, bchi)
    st1 = single.project(1e-11, 1e-14)
    it1 = st1.iters
    phi1, g1, v1 = single.getPhi().copy(), single.getGradPhi().copy(), single.vel.copy()
    single.close()
    ip = npj.IncfloProjection(n, dx, bclo, bchi, opts=npj.nodal_proj_opts(tile=TILE))
    to = lambda a: torch.from_numpy(a).cuda()
    mv = npj.MultiFab.split(vel, n, mg, ng, 3, to=to)
    ms = npj.MultiFab.split(sigma, n, mg, 0, 1, to=to)
    mphi = npj.MultiFab.split(np.zeros((n[2] + 1, n[1] + 1, n[0] + 1)), n, mg, 0, 1, nodal=True, to=to)
    mg_ = npj.MultiFab.split(np.zeros((3, n[2], n[1], n[0])), n, mg, 0, 3, to=to)
    assert len(mv.boxes) == 3 * 2 * 2
    st2 = ip.project_mf(mv, ms, 1.0, mphi, mg_)
    assert st2.iters == it1
    assert np.array_equal(mv.assemble(n), v1[:, 1:-1, 1:-1, 1:-1])
    assert np.array_equal(mphi.assemble(n)[0], phi1) and np.array_equal(mg_.assemble(n), g1)
    # project() never writes ghost cells
    a = mv.arrays[0].cpu().numpy()
    assert np.array_equal(a[:, :, :, 0], vel[:, 0:a.shape[1], 0:a.shape[2], 0])
    ip.close()


def test_bad_box_lists_are_refused():
    import torch
    from incflo_b200 import nodal_projector as npj
    n = (32, 32, 32)
    ip = npj.IncfloProjection(n, (1 / 32,) * 3, (0, 0, 0), (0, 0, 0))
    vel = np.zeros((3, 34, 34, 34)); gp = np.zeros((3, 32, 32, 32)); p = np.zeros((33, 33, 33))
    mv = npj.MultiFab.split(vel, n, 16, 1, 3)
    mgp = npj.MultiFab.split(gp, n, 16, 0, 3)
    mp = npj.MultiFab.split(p, n, 16, 0, 1, nodal=True)
    # a box is missing: the valid boxes no longer tile the domain
    short = npj.MultiFab(mv.boxes[:-1], mv.arrays[:-1], 1, 3)
    with pytest.raises(npj.ProjectionError) as e:
        ip.apply_nodal_projection_mf(short, mgp, mp)
    assert e.value.status == 4
    ip.close()
