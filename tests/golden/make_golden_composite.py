"""Generates tests/golden/composite/*.npz: golden solutions of the COMPOSITE (two AMR level, one fine
box, ratio 2) nodal projection, BASELINE configs[3] scaled down.

Independent of oracle/ and incflo_b200/csrc (no multigrid, no reflux construction): the composite
problem is assembled as ONE Q1 finite-element system on the composite mesh --
  * element stiffness / divergence of every UNCOVERED coarse cell and of every fine cell of the box,
  * unknowns: coarse nodes not strictly inside the box + fine nodes strictly inside the box,
  * fine nodes on the box boundary are hanging nodes: phi1 = trilinear interpolant of the coarse nodes
    of the interface (constraint matrix T), so their element contributions flow into the interface
    coarse equations through T^T (what MLNodeLaplacian::reflux / compRHS do with *_fine_contrib),
-- and solved with a sparse direct solver (Lagrange multiplier for the constant null space).
Then u -= sigma G phi and gphi = G phi per level, average_down of both onto the covered coarse cells
(src/projection/incflo_apply_nodal_projection.cpp:258-266; NodalProjector A.1 (7)) and injection of
phi1 onto the covered coarse nodes.
Run:  python tests/golden/make_golden_composite.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import PER, NEU, DIR, fe_operator, gradT, grad_cells, node_counts, to_full, smooth_random_velocity  # noqa: E402

OUT = os.path.join(HERE, "composite")


def interp1d(nb, periodic=False):
    """linear interpolation of the box's coarse nodes onto its fine nodes in one direction:
    (2nb+1) x (nb+1), or 2nb x nb with wrap when the box spans a periodic direction"""
    if periodic:
        m = sp.lil_matrix((2 * nb, nb))
        for i in range(nb):
            m[2 * i, i] = 1.0
            m[2 * i + 1, i] += 0.5; m[2 * i + 1, (i + 1) % nb] += 0.5
        return m.tocsr()
    m = sp.lil_matrix((2 * nb + 1, nb + 1))
    for i in range(nb + 1):
        m[2 * i, i] = 1.0
    for i in range(nb):
        m[2 * i + 1, i] = 0.5; m[2 * i + 1, i + 1] = 0.5
    return m.tocsr()


def avg_down(f):
    return 0.125 * sum(f[..., c::2, b::2, a::2] for c in range(2) for b in range(2) for a in range(2))


def solve(n0, dx0, bclo, bchi, clo, chi, vel0, ng0, vel1, ng1, sigma0, sigma1, assemble_only=False):
    nb = [chi[d] - clo[d] + 1 for d in range(3)]
    nf = [2 * x for x in nb]
    dx1 = [0.5 * x for x in dx0]
    H3, h3 = float(np.prod(dx0)), float(np.prod(dx1))
    # per direction: the box spans a periodic direction / touches a wall / has a coarse-fine interface
    span = [bclo[d] == PER and clo[d] == 0 and chi[d] == n0[d] - 1 for d in range(3)]
    cf_lo = [not span[d] and clo[d] > 0 for d in range(3)]
    cf_hi = [not span[d] and chi[d] < n0[d] - 1 for d in range(3)]
    for d in range(3):
        assert span[d] or bclo[d] != PER or (clo[d] > 0 and chi[d] < n0[d] - 1)
        assert bclo[d] in (PER, NEU, DIR) and bchi[d] in (PER, NEU, DIR), "golden cases: periodic, wall or outflow faces"
    bc1 = tuple(PER if span[d] else NEU for d in range(3))     # natural (one-sided) sums on every non-periodic face
    cbox = (slice(clo[2], chi[2] + 1), slice(clo[1], chi[1] + 1), slice(clo[0], chi[0] + 1))
    s0 = sigma0.copy()
    s0z = s0.copy(); s0z[cbox] = 0.0
    u0 = vel0[:, ng0:ng0 + n0[2], ng0:ng0 + n0[1], ng0:ng0 + n0[0]].copy()
    u0z = u0.copy(); u0z[(slice(None),) + cbox] = 0.0
    u1 = vel1[:, ng1:ng1 + nf[2], ng1:ng1 + nf[1], ng1:ng1 + nf[0]].copy()
    # element sums, natural (no boundary scaling): K phi = f  <=>  -(sigma grad phi, grad N) = -(u, grad N)
    K0, nn0 = fe_operator(n0, dx0, bclo, bchi, s0z)
    K0 = K0 * H3
    f0 = -H3 * gradT(u0z, n0, dx0, bclo).ravel()
    K1, nn1 = fe_operator(nf, dx1, bc1, bc1, sigma1)
    K1 = K1 * h3
    f1 = -h3 * gradT(u1, nf, dx1, bc1).ravel()
    N0, N1 = K0.shape[0], K1.shape[0]
    # fine nodes on a coarse/fine interface are hanging nodes
    kk, jj, ii = np.meshgrid(np.arange(nn1[2]), np.arange(nn1[1]), np.arange(nn1[0]), indexing="ij")
    idx = (ii, jj, kk)
    bnd = np.zeros(kk.shape, dtype=bool)
    for d in range(3):
        if cf_lo[d]: bnd |= idx[d] == 0
        if cf_hi[d]: bnd |= idx[d] == nf[d]
    bnd = bnd.ravel()
    int_ids = np.flatnonzero(~bnd)
    Ni = int_ids.size
    # trilinear interpolation box coarse nodes -> fine nodes, and box coarse nodes -> global coarse ids
    I1 = [interp1d(nb[d], span[d]) for d in range(3)]
    Bbox = sp.kron(I1[2], sp.kron(I1[1], I1[0])).tocsr()
    rng = [np.arange(n0[d]) if span[d] else np.arange(clo[d], chi[d] + 2) for d in range(3)]
    ck, cj, ci = np.meshgrid(rng[2], rng[1], rng[0], indexing="ij")
    gid = ((ck * nn0[1] + cj) * nn0[0] + ci).ravel()
    S = sp.csr_matrix((np.ones(gid.size), (np.arange(gid.size), gid)), shape=(gid.size, N0))
    Mb = sp.diags(bnd.astype(float))
    E = sp.csr_matrix((np.ones(Ni), (int_ids, np.arange(Ni))), shape=(N1, Ni))
    T = sp.hstack([Mb @ Bbox @ S, E]).tocsr()            # phi1_all = T [phi0; phi1_int]
    A = sp.bmat([[K0, None], [None, sp.csr_matrix((Ni, Ni))]]).tocsr() + T.T @ K1 @ T
    b = np.concatenate([f0, np.zeros(Ni)]) + T.T @ f1
    # active unknowns: coarse nodes not strictly covered, fine nodes that are not hanging
    cin = np.zeros(nn0[::-1], dtype=bool)
    sl = tuple(slice(clo[d] + (1 if cf_lo[d] else 0), (chi[d] + 1) if cf_hi[d] else (n0[d] if span[d] else chi[d] + 2)) for d in (2, 1, 0))
    cin[sl] = True
    # Dirichlet (outflow) faces of the domain: phi = 0 there, on the coarse level and -- where the box touches
    # the face -- on the fine level
    dir0 = np.zeros(nn0[::-1], dtype=bool); dir1 = np.zeros(nn1[::-1], dtype=bool)
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        lo_ = [slice(None)] * 3; hi_ = [slice(None)] * 3
        lo_[ax] = 0; hi_[ax] = -1
        if bclo[d] == DIR:
            dir0[tuple(lo_)] = True
            if clo[d] == 0: dir1[tuple(lo_)] = True
        if bchi[d] == DIR:
            dir0[tuple(hi_)] = True
            if chi[d] == n0[d] - 1: dir1[tuple(hi_)] = True
    act = np.concatenate([~cin.ravel() & ~dir0.ravel(), ~dir1.ravel()[int_ids]])
    inact_free = ~act & ~np.concatenate([dir0.ravel(), dir1.ravel()[int_ids]])   # strictly covered coarse nodes
    assert abs(A[inact_free]).sum() == 0.0 and np.abs(b[inact_free]).max() == 0.0
    Aa = A[act][:, act].tocsc()
    ba = b[act]
    if assemble_only:
        return dict(A=Aa, b=ba, act=act, N0=N0, int_ids=int_ids, nn0=nn0, nn1=nn1, H3=H3, h3=h3)
    if all(x != DIR for x in tuple(bclo) + tuple(bchi)):
        print(f"   compatibility: sum(b) = {ba.sum():.3e} (|b|max {np.abs(ba).max():.3e})")
        c = sp.csc_matrix(np.ones((ba.size, 1)))
        Ksys = sp.bmat([[Aa, c], [c.T, None]]).tocsc()
        x = spl.spsolve(Ksys, np.concatenate([ba, [0.0]]))[:-1]
    else:
        x = spl.spsolve(Aa, ba)
    full = np.zeros(N0 + Ni); full[act] = x
    phi0_u = full[:N0].reshape(nn0[::-1])
    phi1_u = (T @ full).reshape(nn1[::-1])
    nsl = tuple(slice(0, n0[d]) if span[d] else slice(clo[d], chi[d] + 2) for d in (2, 1, 0))
    phi0_u[nsl] = phi1_u[::2, ::2, ::2]                  # injection (interface nodes: identical)
    phi0 = to_full(phi0_u, n0, bclo)
    phi1 = to_full(phi1_u, nf, bc1)
    g1 = grad_cells(phi1, dx1)
    g0 = grad_cells(phi0, dx0)
    s0[cbox] = avg_down(sigma1)
    v0 = vel0.copy(); v1 = vel1.copy()
    v1[:, ng1:ng1 + nf[2], ng1:ng1 + nf[1], ng1:ng1 + nf[0]] = u1 - sigma1[None] * g1
    u0n = u0 - s0[None] * g0
    u0n[(slice(None),) + cbox] = avg_down(v1[:, ng1:ng1 + nf[2], ng1:ng1 + nf[1], ng1:ng1 + nf[0]])
    v0[:, ng0:ng0 + n0[2], ng0:ng0 + n0[1], ng0:ng0 + n0[0]] = u0n
    g0[(slice(None),) + cbox] = avg_down(g1)
    return dict(phi0=phi0, phi1=phi1, gphi0=g0, gphi1=g1, vel0_out=v0, vel1_out=v1)


def cases():
    out = []
    # 1. bouss_bubble-like (BASELINE configs[3] scaled down): periodic x/y, walls z, constant sigma, central box
    n0, dx0 = (16, 16, 16), (1 / 16,) * 3
    bclo = bchi = (PER, PER, NEU)
    clo, chi = (4, 4, 4), (11, 11, 11)
    nf = tuple(2 * (chi[d] - clo[d] + 1) for d in range(3))
    out.append(dict(name="bubble16_box8_const", n0=n0, dx0=dx0, bclo=bclo, bchi=bchi, clo=clo, chi=chi, ng0=2, ng1=2,
                    vel0=smooth_random_velocity(n0, 2, bclo, bchi, 51), vel1=smooth_random_velocity(nf, 2, (NEU,) * 3, (NEU,) * 3, 52),
                    sigma0=np.full(n0[::-1], 0.45 / 16), sigma1=np.full(nf[::-1], 0.45 / 16), var=False))
    # 2. variable density, off-centre non-cubic box, periodic x / walls y,z
    n0, dx0 = (16, 12, 8), (1 / 16,) * 3
    bclo = bchi = (PER, NEU, NEU)
    clo, chi = (3, 2, 2), (10, 7, 5)
    nf = tuple(2 * (chi[d] - clo[d] + 1) for d in range(3))
    rng = np.random.default_rng(7)
    out.append(dict(name="walls_box_offcentre_var", n0=n0, dx0=dx0, bclo=bclo, bchi=bchi, clo=clo, chi=chi, ng0=1, ng1=1,
                    vel0=smooth_random_velocity(n0, 1, bclo, bchi, 61), vel1=smooth_random_velocity(nf, 1, (NEU,) * 3, (NEU,) * 3, 62),
                    sigma0=rng.uniform(0.5, 2.0, size=n0[::-1]), sigma1=rng.uniform(0.5, 2.0, size=nf[::-1]), var=True))
    # 3. rayleigh_taylor-like refinement of the interface region: the fine level spans the periodic x, y
    #    directions completely (a refined slab), walls in z, variable density
    n0, dx0 = (16, 16, 16), (1 / 16,) * 3
    bclo = bchi = (PER, PER, NEU)
    clo, chi = (0, 0, 5), (15, 15, 10)
    nf = tuple(2 * (chi[d] - clo[d] + 1) for d in range(3))
    bc1 = (PER, PER, NEU)
    rng = np.random.default_rng(8)
    out.append(dict(name="rt_slab_periodic_span_var", n0=n0, dx0=dx0, bclo=bclo, bchi=bchi, clo=clo, chi=chi, ng0=1, ng1=1,
                    vel0=smooth_random_velocity(n0, 1, bclo, bchi, 71), vel1=smooth_random_velocity(nf, 1, bc1, bc1, 72),
                    sigma0=rng.uniform(0.5, 2.0, size=n0[::-1]), sigma1=rng.uniform(0.5, 2.0, size=nf[::-1]), var=True))
    # 4. closed box, fine box in a corner: touches the x-lo and z-hi walls, interface on the other four faces
    n0, dx0 = (12, 12, 12), (1 / 12,) * 3
    bclo = bchi = (NEU, NEU, NEU)
    clo, chi = (0, 3, 6), (5, 8, 11)
    nf = tuple(2 * (chi[d] - clo[d] + 1) for d in range(3))
    out.append(dict(name="closed_box_corner_walls_const", n0=n0, dx0=dx0, bclo=bclo, bchi=bchi, clo=clo, chi=chi, ng0=2, ng1=1,
                    vel0=smooth_random_velocity(n0, 2, bclo, bchi, 81), vel1=smooth_random_velocity(nf, 1, bclo, bchi, 82),
                    sigma0=np.full(n0[::-1], 0.8), sigma1=np.full(nf[::-1], 0.8), var=False))
    return out


def cases_dirichlet():
    """non-singular problems (an outflow face): tests/golden/composite_dirichlet/"""
    out = []
    # 1. channel-like: wall x-lo, outflow x-hi, walls y, periodic z; interior box
    n0, dx0 = (16, 8, 8), (1 / 16,) * 3
    bclo, bchi = (NEU, NEU, PER), (DIR, NEU, PER)
    clo, chi = (4, 2, 2), (9, 5, 5)
    nf = tuple(2 * (chi[d] - clo[d] + 1) for d in range(3))
    nat = (NEU,) * 3
    rng = np.random.default_rng(9)
    out.append(dict(name="channel_outflow_interior_box_var", n0=n0, dx0=dx0, bclo=bclo, bchi=bchi, clo=clo, chi=chi, ng0=1, ng1=1,
                    vel0=smooth_random_velocity(n0, 1, nat, nat, 91), vel1=smooth_random_velocity(nf, 1, nat, nat, 92),
                    sigma0=rng.uniform(0.5, 2.0, size=n0[::-1]), sigma1=rng.uniform(0.5, 2.0, size=nf[::-1]), var=True))
    # 2. the fine box sits ON the outflow face (and spans the periodic direction)
    clo, chi = (10, 2, 0), (15, 5, 7)
    nf = tuple(2 * (chi[d] - clo[d] + 1) for d in range(3))
    out.append(dict(name="channel_box_on_outflow_face_const", n0=n0, dx0=dx0, bclo=bclo, bchi=bchi, clo=clo, chi=chi, ng0=1, ng1=1,
                    vel0=smooth_random_velocity(n0, 1, nat, nat, 93), vel1=smooth_random_velocity(nf, 1, nat, nat, 94),
                    sigma0=np.full(n0[::-1], 0.6), sigma1=np.full(nf[::-1], 0.6), var=False))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(OUT + "_dirichlet", exist_ok=True)
    for c in cases() + cases_dirichlet():
        print(c["name"])
        r = solve(c["n0"], c["dx0"], c["bclo"], c["bchi"], c["clo"], c["chi"], c["vel0"], c["ng0"], c["vel1"], c["ng1"],
                  c["sigma0"], c["sigma1"])
        path = os.path.join(OUT + ("_dirichlet" if DIR in tuple(c["bclo"]) + tuple(c["bchi"]) else ""), c["name"] + ".npz")
        np.savez_compressed(path, n0=np.array(c["n0"]), dx0=np.array(c["dx0"]), bclo=np.array(c["bclo"]), bchi=np.array(c["bchi"]),
                            clo=np.array(c["clo"]), chi=np.array(c["chi"]), ng0=c["ng0"], ng1=c["ng1"], vel0_in=c["vel0"],
                            vel1_in=c["vel1"], sigma0=c["sigma0"], sigma1=c["sigma1"], var=c["var"], **r)
        print(f"   |phi1|max={np.abs(r['phi1']).max():.4e} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
