"""Generates the golden fixtures tests/golden/*.npz for the nodal projection.

The reference ships no golden vectors for this path and its arithmetic (AMReX MLMG) cannot be
built offline, so these fixtures come from an INDEPENDENT evaluation of the same
discrete problem -- no multigrid, no code shared with oracle/ or incflo_b200/csrc:
  * L   : Q1 finite-element stiffness assembled element by element (scipy.sparse), rows divided by
          the node weight at reflecting boundaries (AMReX's reflected operator)
  * G   : cell gradient from the 8 corner nodes (A.7), written with numpy slices
  * D   : nodal divergence as the negative adjoint of G, D = -W^-1 G^T, plus the inflow ghost-cell
          flux of A.2 written out explicitly
  * phi : sparse direct solve of L phi = D u (Lagrange multiplier on the weighted mean when singular)
  * u  <- u - sigma G phi,   gphi = G phi
Run:  python tests/golden/make_golden.py      (seconds; deterministic)
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

PER, NEU, DIR, INF = 0, 1, 2, 3


def node_counts(n, bclo):
    return [n[d] + (0 if bclo[d] == PER else 1) for d in range(3)]


def fe_operator(n, dx, bclo, bchi, sigma):
    """-(sigma grad N_a, grad N_b)/cell volume on unique nodes, as a sparse matrix (vectorised assembly)"""
    nn = node_counts(n, bclo)
    K1 = lambda h: np.array([[1.0, -1.0], [-1.0, 1.0]]) / h
    M1 = lambda h: np.array([[2.0, 1.0], [1.0, 2.0]]) * h / 6.0
    Ke = (np.einsum("ad,be,cf->abcdef", K1(dx[0]), M1(dx[1]), M1(dx[2])) +
          np.einsum("ad,be,cf->abcdef", M1(dx[0]), K1(dx[1]), M1(dx[2])) +
          np.einsum("ad,be,cf->abcdef", M1(dx[0]), M1(dx[1]), K1(dx[2]))) / (dx[0] * dx[1] * dx[2])
    k, j, i = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")

    def nid(ii, jj, kk):
        ii = ii % n[0] if bclo[0] == PER else ii
        jj = jj % n[1] if bclo[1] == PER else jj
        kk = kk % n[2] if bclo[2] == PER else kk
        return ((kk * nn[1] + jj) * nn[0] + ii).ravel()
    rows, cols, vals = [], [], []
    for a in range(2):
        for b in range(2):
            for c in range(2):
                for d in range(2):
                    for e in range(2):
                        for f in range(2):
                            rows.append(nid(i + a, j + b, k + c)); cols.append(nid(i + d, j + e, k + f))
                            vals.append((-sigma * Ke[a, b, c, d, e, f]).ravel())
    N = int(np.prod(nn))
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N)), nn


def node_weights(n, bclo, bchi):
    """dot-product weight: 0 on Dirichlet nodes, 1/2 per reflecting (Neumann/inflow) face"""
    nn = node_counts(n, bclo)
    w = np.ones((nn[2], nn[1], nn[0]))
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        for side, bc in ((0, bclo[d]), (1, bchi[d])):
            sl = [slice(None)] * 3
            sl[ax] = 0 if side == 0 else -1
            if bc in (NEU, INF):
                w[tuple(sl)] *= 0.5
            elif bc == DIR:
                w[tuple(sl)] = 0.0
    return w


def grad_cells(phi_full, dx):
    """G phi on cells from the (n+1)^3 nodal array (A.7)"""
    p = phi_full
    gx = (p[:-1, :-1, 1:] - p[:-1, :-1, :-1] + p[:-1, 1:, 1:] - p[:-1, 1:, :-1] +
          p[1:, :-1, 1:] - p[1:, :-1, :-1] + p[1:, 1:, 1:] - p[1:, 1:, :-1]) * (0.25 / dx[0])
    gy = (p[:-1, 1:, :-1] - p[:-1, :-1, :-1] + p[:-1, 1:, 1:] - p[:-1, :-1, 1:] +
          p[1:, 1:, :-1] - p[1:, :-1, :-1] + p[1:, 1:, 1:] - p[1:, :-1, 1:]) * (0.25 / dx[1])
    gz = (p[1:, :-1, :-1] - p[:-1, :-1, :-1] + p[1:, :-1, 1:] - p[:-1, :-1, 1:] +
          p[1:, 1:, :-1] - p[:-1, 1:, :-1] + p[1:, 1:, 1:] - p[:-1, 1:, 1:]) * (0.25 / dx[2])
    return np.stack([gx, gy, gz])


def to_full(phi_u, n, bclo):
    """unique-node array -> (n+1)^3 with the periodic image planes appended"""
    p = phi_u
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        if bclo[d] == PER:
            p = np.concatenate([p, np.take(p, [0], axis=ax)], axis=ax)
    return p


def gradT(u, n, dx, bclo):
    """G^T u: scatter of the cell values onto the unique nodes"""
    nn = node_counts(n, bclo)
    full = np.zeros((n[2] + 1, n[1] + 1, n[0] + 1))
    for comp, (h, ax) in enumerate(((dx[0], 2), (dx[1], 1), (dx[2], 0))):
        for c in range(2):
            for b in range(2):
                for a in range(2):
                    sgn = (a, b, c)[comp] * 2 - 1
                    full[c:c + n[2], b:b + n[1], a:a + n[0]] += sgn * (0.25 / h) * u[comp]
    for d, ax in ((0, 2), (1, 1), (2, 0)):   # fold the periodic image plane back
        if bclo[d] == PER:
            first = [slice(None)] * 3; last = [slice(None)] * 3
            first[ax] = 0; last[ax] = -1
            full[tuple(first)] += full[tuple(last)]
            full = np.delete(full, -1, axis=ax)
    assert list(full.shape) == nn[::-1]
    return full


def divergence(vel, ng, n, dx, bclo, bchi):
    """rhs = D u (A.2) = -W^-1 G^T u_valid  +  inflow ghost flux"""
    w = node_weights(n, bclo, bchi)
    u = vel[:, ng:ng + n[2], ng:ng + n[1], ng:ng + n[0]]
    rhs = -gradT(u, n, dx, bclo)
    # inflow faces: the normal ghost velocity enters like one more cell layer (tangential ghost
    # velocity is never seen); scatter -(+-1) * u_n / (4 h) of the ghost layer onto the face nodes
    nn = node_counts(n, bclo)
    for d, ax in ((0, 3), (1, 2), (2, 1)):
        for side, bc in ((0, bclo[d]), (1, bchi[d])):
            if bc != INF:
                continue
            sl = [slice(None), slice(ng, ng + n[2]), slice(ng, ng + n[1]), slice(ng, ng + n[0])]
            sl[ax] = ng - 1 if side == 0 else ng + n[d]
            un = vel[tuple(sl)][d]                      # normal component on the ghost layer, 2-D (the other two axes)
            sgn = 1.0 if side == 0 else -1.0           # ghost cell sits on the -/+ side of the face node
            face = np.zeros([s + 1 for s in un.shape])
            for b in range(2):
                for a in range(2):
                    face[b:b + un.shape[0], a:a + un.shape[1]] += -sgn * (0.25 / dx[d]) * un
            # fold periodic tangential directions
            tang = [t for t in (2, 1, 0) if t != d]    # axes of `un` in (slow, fast) order are the remaining of (z,y,x)
            tang = [t for t in (2, 1, 0) if t != d]
            for axis_pos, t in enumerate(sorted(tang, reverse=True)):
                if bclo[t] == PER:
                    first = [slice(None)] * 2; last = [slice(None)] * 2
                    first[axis_pos] = 0; last[axis_pos] = -1
                    face[tuple(first)] += face[tuple(last)]
                    face = np.delete(face, -1, axis=axis_pos)
            idx = [slice(None)] * 3
            idx[{0: 2, 1: 1, 2: 0}[d]] = 0 if side == 0 else nn[d] - 1
            rhs[tuple(idx)] += face
    rhs = np.where(w > 0, rhs / np.where(w > 0, w, 1.0), 0.0)
    return rhs, w


def project(n, dx, bclo, bchi, vel, ng, sigma, dmask=None):
    """dmask: optional boolean unique-node array of extra Dirichlet nodes (the zeros of an overset mask)"""
    sig = sigma if isinstance(sigma, np.ndarray) else np.full((n[2], n[1], n[0]), float(sigma))
    A, nn = fe_operator(n, dx, bclo, bchi, sig)
    rhs, w = divergence(vel, ng, n, dx, bclo, bchi)
    if dmask is not None:
        w = np.where(dmask, 0.0, w)
        rhs = np.where(dmask, 0.0, rhs)
    m = (w > 0).ravel()
    singular = all(b != DIR for b in tuple(bclo) + tuple(bchi)) and (dmask is None or not dmask.any())
    Wi = sp.diags(1.0 / np.where(m, w.ravel(), 1.0))
    Ar = (Wi @ A)[m][:, m].tocsc()
    b = rhs.ravel()[m].copy()
    if singular:
        wm = w.ravel()[m]
        b -= (wm * b).sum() / wm.sum()
        c = sp.csc_matrix(wm[:, None])
        K = sp.bmat([[Ar, c], [c.T, None]]).tocsc()
        x = spl.spsolve(K, np.concatenate([b, [0.0]]))[:-1]
    else:
        x = spl.spsolve(Ar, b)
    phi_u = np.zeros(w.size); phi_u[m] = x
    phi_u = phi_u.reshape(w.shape)
    phi = to_full(phi_u, n, bclo)
    g = grad_cells(phi, dx)
    out = vel.copy()
    out[:, ng:ng + n[2], ng:ng + n[1], ng:ng + n[0]] -= sig[None] * g
    return dict(phi=phi, gphi=g, vel_out=out, rhs=to_full(rhs, n, bclo))


def smooth_random_velocity(n, ng, bclo, bchi, seed, inflow=0.3):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((3, n[2], n[1], n[0]))
    for ax in (1, 2, 3):
        v = 0.5 * v + 0.25 * (np.roll(v, 1, ax) + np.roll(v, -1, ax))
    vel = np.zeros((3, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng))
    vel[:, ng:ng + n[2], ng:ng + n[1], ng:ng + n[0]] = v
    for d, ax in ((0, 3), (1, 2), (2, 1)):
        for side, bc in ((0, bclo[d]), (1, bchi[d])):
            if bc == INF:
                sl = [slice(None)] * 4
                sl[ax] = ng - 1 if side == 0 else ng + n[d]
                vel[tuple(sl)] = inflow + 0.1 * rng.standard_normal(vel[tuple(sl)].shape)
    return vel


def cases():
    from incflo_b200 import problems
    out = []
    # 1. test_no_eb_3d/benchmark.taylor_green_vortices-like: periodic, constant sigma (BASELINE configs[0] scaled down)
    cfg = problems.make("tgv", 16, ng=1, device="cpu")
    out.append(dict(name="tgv16_periodic_const", n=cfg["n"], dx=cfg["dx"], bclo=cfg["bclo"], bchi=cfg["bchi"],
                    vel=cfg["vel"].numpy().copy(), ng=1, sigma=float(cfg["const_sigma"])))
    # 2. rayleigh_taylor-like: periodic x/y, walls z, sigma = dt/rho (BASELINE configs[1] scaled down)
    cfg = problems.make("rt", 16, ng=1, device="cpu")
    out.append(dict(name="rt16_walls_var", n=cfg["n"], dx=cfg["dx"], bclo=cfg["bclo"], bchi=cfg["bchi"],
                    vel=cfg["vel"].numpy().copy(), ng=1, sigma=cfg["sigma"].numpy().copy()))
    # 3. channel: mass inflow xlo, pressure outflow xhi, no-slip walls y, periodic z, random sigma
    n, dx = (24, 8, 8), (1 / 24,) * 3
    bclo, bchi = (INF, NEU, PER), (DIR, NEU, PER)
    rng = np.random.default_rng(3)
    out.append(dict(name="channel_inflow_outflow_var", n=n, dx=dx, bclo=bclo, bchi=bchi,
                    vel=smooth_random_velocity(n, 2, bclo, bchi, 31), ng=2, sigma=rng.uniform(0.5, 2.0, size=(n[2], n[1], n[0]))))
    # 4. closed box, anisotropic cells, constant sigma
    n, dx = (12, 8, 10), (0.1, 0.07, 0.05)
    bclo = bchi = (NEU, NEU, NEU)
    out.append(dict(name="box_neumann_aniso_const", n=n, dx=dx, bclo=bclo, bchi=bchi,
                    vel=smooth_random_velocity(n, 1, bclo, bchi, 41), ng=1, sigma=0.37))
    # 5. mixed BC (incflo probtype 1101-like): x faces are "mixed" = LinOpBCType::inflow + overset mask, split along y;
    #    x-lo: nodes j <= ny/2 are outflow (Dirichlet), x-hi: nodes j > ny/2 (make_nodalBC_mask / prob_set_BC_MF)
    n, dx = (16, 12, 8), (1 / 16,) * 3
    bclo, bchi = (INF, NEU, NEU), (INF, NEU, NEU)
    half = n[1] // 2
    kk, jj, ii = np.meshgrid(np.arange(n[2] + 1), np.arange(n[1] + 1), np.arange(n[0] + 1), indexing="ij")
    dmask = ((ii == 0) & (jj <= half)) | ((ii == n[0]) & (jj > half))
    rng = np.random.default_rng(5)
    out.append(dict(name="mixed_x_split_y_var", n=n, dx=dx, bclo=bclo, bchi=bchi, vel=smooth_random_velocity(n, 1, bclo, bchi, 51), ng=1,
                    sigma=rng.uniform(0.5, 2.0, size=(n[2], n[1], n[0])), dmask=dmask, mixed=dict(lo=(1, 0, 0), hi=(1, 0, 0), dir=1, half=half)))
    return out


def main():
    for c in cases():
        r = project(c["n"], c["dx"], c["bclo"], c["bchi"], c["vel"], c["ng"], c["sigma"], c.get("dmask"))
        var = isinstance(c["sigma"], np.ndarray)
        path = os.path.join(HERE, c["name"] + ".npz")
        np.savez_compressed(path, n=np.array(c["n"]), dx=np.array(c["dx"]), bclo=np.array(c["bclo"]), bchi=np.array(c["bchi"]),
                            ng=c["ng"], vel_in=c["vel"], sigma=c["sigma"] if var else np.array(c["sigma"]), var=var,
                            phi=r["phi"], gphi=r["gphi"], vel_out=r["vel_out"], rhs=r["rhs"],
                            **({} if "mixed" not in c else dict(mixed_lo=np.array(c["mixed"]["lo"]), mixed_hi=np.array(c["mixed"]["hi"]),
                                                                mix_dir=c["mixed"]["dir"], mix_half=c["mixed"]["half"])))
        print(f"{c['name']}: n={c['n']} |phi|max={np.abs(r['phi']).max():.4e} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
