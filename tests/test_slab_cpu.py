"""CPU coverage of the N > 1 host logic: the slab partition exported by the C ABI and the halo
exchange plan, exercised with two gloo processes (no GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_slab_partition_covers_domain():
    from incflo_b200 import slab
    for per in (True, False):
        bclo = (0, 0, 0 if per else 1)
        for nz, P in ((64, 1), (64, 2), (256, 8), (96, 4)):
            cells, nodes = [], []
            for r in range(P):
                clo, chi, nlo, nhi = slab.slab_range((16, 16, nz), bclo, r, P)
                cells += list(range(clo, chi + 1)); nodes += list(range(nlo, nhi + 1))
            assert cells == list(range(nz))
            assert nodes == list(range(nz if per else nz + 1))  # every unique node plane owned exactly once
    with pytest.raises(Exception):
        slab.slab_range((16, 16, 30), (0, 0, 0), 0, 4)


def test_distributed_level_plan():
    """agglomeration plan of the C ABI (b200np_dist_plan): which levels stay slab-distributed"""
    from incflo_b200 import nodal_projector as npj, slab
    # bench weak scaling, 256^3 per GPU: levels with 256, 128, 64 planes per rank are distributed
    assert slab.distributed_levels((256, 256, 512), 2) == (3, 8)
    assert slab.distributed_levels((256, 256, 2048), 8) == (3, 8)
    # strong scaling 512^3 / 1024^3 on 8 GPUs: big levels with thin slabs must NOT be replicated
    assert slab.distributed_levels((512, 512, 512), 8)[0] == 2      # 64, 32 (256^3 nodes: too big to replicate)
    assert slab.distributed_levels((1024, 1024, 1024), 8)[0] == 3   # 128, 64, 32
    assert slab.distributed_levels((1024, 1024, 1024), 4)[0] == 3   # 256, 128, 64; 128^3 with 32 planes is replicated
    # explicit threshold (the slab parity tests run with 8): every level down to 8 planes per rank
    assert slab.distributed_levels((64, 64, 64), 2, min_planes=8)[0] == 3
    assert slab.distributed_levels((64, 64, 64), 4, min_planes=8)[0] == 2
    assert slab.distributed_levels((64, 64, 64), 2)[0] == 1         # default: only level 0
    assert slab.distributed_levels((64, 64, 64), 1) == (0, 6)
    with pytest.raises(npj.ProjectionError):                        # level 0 cannot be cut into even slabs
        slab.distributed_levels((64, 64, 60), 8)


def _worker(rank, world, port, periodic, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from incflo_b200 import slab
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nz, ny, nx = 16, 3, 5
    bclo = (0, 0, 0 if periodic else 1)
    rng = np.random.default_rng(5)
    nn = nz if periodic else nz + 1
    glob = rng.standard_normal((nn, ny, nx))          # unique node planes of the global field
    clo, chi, nlo, nhi = slab.slab_range((nx, ny, nz), bclo, rank, world)
    own = torch.from_numpy(glob[nlo:nhi + 1].copy())
    ghost = {"lo": torch.zeros(ny, nx, dtype=torch.float64), "hi": torch.zeros(ny, nx, dtype=torch.float64)}
    sends, recvs = slab.halo_plan(rank, world, periodic)
    reqs = [dist.isend(own[0 if which == "first" else -1].contiguous(), peer) for peer, which in sends]
    for peer, slot in recvs:
        dist.recv(ghost[slot], peer)
    for rq in reqs:
        rq.wait()
    has = {s for _, s in recvs}
    if "lo" not in has:   # physical end: reflection phi(-1) = phi(1)
        ghost["lo"] = own[1].clone()
    if "hi" not in has:
        ghost["hi"] = own[-2].clone()
    want_lo = glob[(nlo - 1) % nn] if periodic else glob[abs(nlo - 1)]
    k = nhi + 1
    want_hi = glob[k % nn] if periodic else glob[k if k <= nz else 2 * nz - k]
    ok = np.array_equal(ghost["lo"].numpy(), want_lo) and np.array_equal(ghost["hi"].numpy(), want_hi)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("periodic", [True, False])
def test_halo_plan_two_gloo_ranks(periodic):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + (1 if periodic else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, periodic, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
