"""Host-side description of the z-slab decomposition used by libb200np.so.

Mirrors, in Python, the partition (b200np_slab_range) and the neighbour exchange plan
(exchange_planes in csrc/b200np.cu) so that the multi-process logic can be exercised on CPU with
torch.distributed's gloo backend and so that bench.py / tests can cut global fields into slabs.
"""
from . import nodal_projector as npj


def halo_plan(rank, nranks, periodic_z):
    """Order of the grouped point-to-point operations of one plane exchange.

    Returns (sends, recvs): sends = [(peer, 'first'|'last')], recvs = [(peer, 'hi'|'lo')] -- the
    first owned plane goes to the lower neighbour's upper ghost slot, the last owned plane to the
    upper neighbour's lower ghost slot.  Ends of a non-periodic domain have no partner (the ghost
    slot is filled locally: reflection for nodes, clamp for cells)."""
    has_lo = periodic_z or rank > 0
    has_hi = periodic_z or rank < nranks - 1
    lo, hi = (rank - 1) % nranks, (rank + 1) % nranks
    sends, recvs = [], []
    if has_lo:
        sends.append((lo, "first"))
    if has_hi:
        sends.append((hi, "last"))
    if has_hi:
        recvs.append((hi, "hi"))
    if has_lo:
        recvs.append((lo, "lo"))
    return sends, recvs


def distributed_levels(n_cell, nranks, min_planes=0, bclo=(0, 0, 0), max_coarsening_level=100):
    """(levels that stay slab-distributed, levels in total) for the domain n_cell = (nx, ny, nz) on nranks
    ranks (b200np_dist_plan; needs no GPU).  Level 0 is always distributed; a coarser level stays distributed
    while every rank keeps >= min_planes (default 64) cell planes or while it is too big to replicate
    (> 128^3 nodes, down to 8 planes per rank); the rest is replicated on every rank (agglomeration)."""
    import ctypes as C
    from . import _lib
    g = _lib.Geom()
    for d in range(3):
        g.n_cell[d] = int(n_cell[d]); g.dx[d] = 1.0; g.bc_lo[d] = int(bclo[d]); g.bc_hi[d] = int(bclo[d])
    nd, nl = C.c_int(), C.c_int()
    rc = _lib.lib().b200np_dist_plan(C.byref(g), int(nranks), int(min_planes), int(max_coarsening_level), C.byref(nd), C.byref(nl))
    if rc != 0:
        raise npj.ProjectionError(rc)
    return nd.value, nl.value


def cut(global_arr, zlo, zhi, ng, node=False):
    """local box [zlo-ng, zhi+ng] (cells) or [zlo, zhi+1] (nodes, ng must be 0) of a global array whose
    z axis is axis -3 and already carries ng ghost planes."""
    if node:
        return global_arr[..., zlo:zhi + 2, :, :].copy()
    return global_arr[..., zlo:zhi + 1 + 2 * ng, :, :].copy()


slab_range = npj.slab_range
