"""Golden fixtures tests/golden/eb_*.npz for the EB nodal projection (MLNodeLaplacian with an EB factory, BASELINE configs[4]).

Independent evaluation of the discrete problem -- no multigrid, no stencils, no monomial integrals, no code shared with
oracle/eb_oracle.py or incflo_b200/csrc:
  * the fluid part of every cut cell is the polyhedron (unit cell) n (half-space of the cell's EB plane), split into tetrahedra;
  * the element matrices  K_c[a][b] = sum_d dxinv_d^2 int_F d_d N_a d_d N_b,  the gradient integrals  int_F d_d N_a,  the volume
    and the EB-face integrals  int_EB N_a dA  are evaluated by Gauss quadrature of the trilinear shape functions THEMSELVES over
    those tetrahedra / triangles (uncut cells: tensor Gauss rule on the cube);
  * the global matrix is assembled element by element into scipy.sparse, Dirichlet and covered nodes are eliminated, the system is
    solved directly (mean-free right-hand side + Lagrange multiplier when no face is Dirichlet);
  * rhs = -sum_c u_c . int_F grad N_a  with the first ghost layer's normal velocity at non-periodic faces (+ EB inflow term),
    u -= sigma * (1/V) int_F grad phi.
The geometry (per-cell plane) and the 18 + 8 monomial integrals stored in the fixture as INPUTS for the oracle / the CUDA path come
from incflo_b200/eb_geometry.py.
Run: python tests/golden/make_golden_eb.py
"""
import itertools
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from incflo_b200 import eb_geometry as eg  # noqa: E402

PER, NEU, DIR, INF = 0, 1, 2, 3
CORN = [(a & 1, (a >> 1) & 1, (a >> 2) & 1) for a in range(8)]
SGN = np.array([[2 * c - 1 for c in ca] for ca in CORN], dtype=np.float64)   # (8, 3)

_g5, _w5 = np.polynomial.legendre.leggauss(5)
_g5, _w5 = 0.5 * (_g5 + 1.0), 0.5 * _w5


def shape(pts):
    """N (nq, 8) and grad N (nq, 8, 3) of the trilinear functions on the unit cell centred at 0"""
    f = 0.5 + SGN[None, :, :] * pts[:, None, :]                  # (nq, 8, 3)
    N = f[..., 0] * f[..., 1] * f[..., 2]
    dN = np.stack([SGN[None, :, 0] * f[..., 1] * f[..., 2], SGN[None, :, 1] * f[..., 0] * f[..., 2],
                   SGN[None, :, 2] * f[..., 0] * f[..., 1]], axis=-1)
    return N, dN


def tet_rule(v):
    r, s, t = np.meshgrid(_g5, _g5, _g5, indexing="ij")
    w = (_w5[:, None, None] * _w5[None, :, None] * _w5[None, None, :] * (1 - r) ** 2 * (1 - s)).ravel()
    l1, l2, l3 = r.ravel(), (s * (1 - r)).ravel(), (t * (1 - r) * (1 - s)).ravel()
    l0 = 1 - l1 - l2 - l3
    pts = l0[:, None] * v[0] + l1[:, None] * v[1] + l2[:, None] * v[2] + l3[:, None] * v[3]
    return pts, w * abs(np.linalg.det(v[1:] - v[0]))


def tri_rule(v):
    r, s = np.meshgrid(_g5, _g5, indexing="ij")
    w = (_w5[:, None] * _w5[None, :] * (1 - r)).ravel()
    l1, l2 = r.ravel(), (s * (1 - r)).ravel()
    l0 = 1 - l1 - l2
    pts = l0[:, None] * v[0] + l1[:, None] * v[1] + l2[:, None] * v[2]
    return pts, w * np.linalg.norm(np.cross(v[1] - v[0], v[2] - v[0]))


def cube_rule():
    g, w = np.polynomial.legendre.leggauss(3)
    g, w = 0.5 * g, 0.5 * w
    pts = np.array(list(itertools.product(g, g, g)))
    ww = np.array([a * b * c for a, b, c in itertools.product(w, w, w)])
    return pts, ww


def cell_integrals(pts, w, dxinv):
    N, dN = shape(pts)
    K = np.einsum("q,qad,qbd,d->ab", w, dN, dN, np.asarray(dxinv) ** 2)
    G = np.einsum("q,qad->ad", w, dN)
    return K, G, w.sum()


def element_data(geom, dxinv):
    """per cell: K (8, 8), G (8, 3), V, BN (8,) = int_EB N_a dA"""
    nx, ny, nz = geom.n
    K = np.zeros((nz, ny, nx, 8, 8)); G = np.zeros((nz, ny, nx, 8, 3)); V = np.zeros((nz, ny, nx)); BN = np.zeros((nz, ny, nx, 8))
    Kr, Gr, Vr = cell_integrals(*cube_rule(), dxinv)
    reg = geom.vfrac == 1.0
    K[reg], G[reg], V[reg] = Kr, Gr, Vr
    cache = {}
    for k, j, i in zip(*np.nonzero(geom.cut_mask())):
        key = (tuple(np.round(geom.bnorm[:, k, j, i], 14)), round(float(geom.boff[k, j, i]), 14))
        if key not in cache:
            tets, tris = eg.cut_cell_simplices(geom.bnorm[:, k, j, i], geom.boff[k, j, i])
            Kc = np.zeros((8, 8)); Gc = np.zeros((8, 3)); Vc = 0.0; Bc = np.zeros(8)
            for t in tets:
                a, b, c = cell_integrals(*tet_rule(t), dxinv)
                Kc += a; Gc += b; Vc += c
            for t in tris:
                pts, w = tri_rule(t)
                Bc += w @ shape(pts)[0]
            cache[key] = (Kc, Gc, Vc, Bc)
        K[k, j, i], G[k, j, i], V[k, j, i], BN[k, j, i] = cache[key]
    return K, G, V, BN


def node_index(n, per):
    nn = [n[d] if per[d] else n[d] + 1 for d in range(3)]
    return nn, np.arange(nn[0] * nn[1] * nn[2]).reshape(nn[2], nn[1], nn[0])


def cell_nodes(n, per, nn, idx, a):
    """global node of corner a for every cell: (nz, ny, nx)"""
    ii = (np.arange(n[0]) + a[0]) % nn[0] if per[0] else np.arange(n[0]) + a[0]
    jj = (np.arange(n[1]) + a[1]) % nn[1] if per[1] else np.arange(n[1]) + a[1]
    kk = (np.arange(n[2]) + a[2]) % nn[2] if per[2] else np.arange(n[2]) + a[2]
    return idx[np.ix_(kk, jj, ii)]


def project(n, dx, bclo, bchi, geom, vel, sigma, eb_vel=None):
    per = [b == PER for b in bclo]
    dxinv = [1.0 / h for h in dx]
    K, G, V, BN = element_data(geom, dxinv)
    nn, idx = node_index(n, per)
    N = idx.size
    sig = np.broadcast_to(np.asarray(sigma, dtype=np.float64), V.shape)
    cn = [cell_nodes(n, per, nn, idx, a) for a in CORN]
    rows, cols, vals = [], [], []
    for a in range(8):
        for b in range(8):
            rows.append(cn[a].ravel()); cols.append(cn[b].ravel()); vals.append((-sig * K[..., a, b]).ravel())
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    # right-hand side: interior cells
    u = vel[:, 1:-1, 1:-1, 1:-1]
    rhs = np.zeros(N)
    for a in range(8):
        c = -(dxinv[0] * u[0] * G[..., a, 0] + dxinv[1] * u[1] * G[..., a, 1] + dxinv[2] * u[2] * G[..., a, 2])
        if eb_vel is not None:
            vn = eb_vel[0] * geom.bnorm[0] + eb_vel[1] * geom.bnorm[1] + eb_vel[2] * geom.bnorm[2]
            c = c + dxinv[0] * vn * BN[..., a]
        np.add.at(rhs, cn[a].ravel(), c.ravel())
    # ghost cells beyond ONE non-periodic face: normal velocity only, geometry of the adjacent interior cell
    for d in range(3):
        if per[d]:
            continue
        ax = 2 - d
        for side in (0, 1):
            # ghost cell index -1 (side 0) or n (side 1); it touches the boundary nodes with its corners a_d = 1 (side 0) / 0 (side 1)
            gsl = [slice(1, -1)] * 3
            gsl[ax] = 0 if side == 0 else n[d] + 1
            ug = vel[(d,) + tuple(gsl)]                          # normal component, 2-D array over the face
            isl = [slice(None)] * 3
            isl[ax] = 0 if side == 0 else n[d] - 1
            for a in range(8):
                if CORN[a][d] != (1 if side == 0 else 0):
                    continue
                # the ghost cell's corner a coincides with the interior cell's corner a' (a'_d flipped); int d_d N_a over the ghost
                # cell with the mirrored geometry = the interior cell's integral for a' with the opposite sign of s_d
                ap = a ^ (1 << d)
                Gg = -G[tuple(isl) + (ap, d)]
                node = cn[ap][tuple(isl)]
                np.add.at(rhs, node.ravel(), (-dxinv[d] * ug * Gg).ravel())
    # boundary conditions and covered nodes
    dm = np.zeros(idx.shape, dtype=bool)
    for d in range(3):
        if per[d]:
            continue
        sl = [slice(None)] * 3
        if bclo[d] == DIR:
            sl[2 - d] = 0; dm[tuple(sl)] = True
        if bchi[d] == DIR:
            sl[2 - d] = nn[d] - 1; dm[tuple(sl)] = True
    diag = A.diagonal()
    active = (~dm.ravel()) & (diag != 0.0)
    ia = np.nonzero(active)[0]
    Aa = A[ia][:, ia].tocsc()
    ra = rhs[ia].copy()
    singular = all(b != DIR for b in tuple(bclo) + tuple(bchi))
    if singular:
        ra -= ra.mean()
        one = sp.csc_matrix(np.ones((ia.size, 1)))
        Kmat = sp.bmat([[Aa, one], [one.T, None]]).tocsc()
        x = spl.spsolve(Kmat, np.concatenate([ra, [0.0]]))[:-1]
    else:
        x = spl.spsolve(Aa, ra)
    phi = np.zeros(N)
    phi[ia] = x
    rhs_full = np.zeros(N)
    rhs_full[ia] = ra
    # update
    g = np.zeros((3,) + V.shape)
    for a in range(8):
        pa = phi[cn[a]]
        for d in range(3):
            g[d] += dxinv[d] * pa * G[..., a, d]
    fluid = V > 0
    g = np.where(fluid, g / np.where(fluid, V, 1.0), 0.0)
    unew = np.where(fluid, u - sig * g, 0.0)
    return dict(phi=phi.reshape(idx.shape), rhs=rhs_full.reshape(idx.shape), gphi=g, vel_new=unew, vfrac_quad=V,
                resid=np.abs(Aa @ x - ra).max())


def smooth_velocity(n, h, rng, base=(1.0, 0.0, 0.0)):
    nx, ny, nz = n
    x = (np.arange(-1, nx + 1) + 0.5) * h
    y = (np.arange(-1, ny + 1) + 0.5) * h
    z = (np.arange(-1, nz + 1) + 0.5) * h
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    lx, ly, lz = nx * h, ny * h, nz * h
    tp = 2 * np.pi
    vel = np.empty((3,) + X.shape)
    ph = rng.uniform(0, tp, size=6)
    vel[0] = base[0] + 0.3 * np.sin(tp * X / lx + ph[0]) * np.cos(tp * Y / ly + ph[1]) * np.cos(tp * Z / lz)
    vel[1] = base[1] + 0.3 * np.cos(tp * X / lx + ph[2]) * np.sin(tp * Y / ly + ph[3]) * np.cos(tp * Z / lz + 0.3)
    vel[2] = base[2] + 0.2 * np.cos(tp * X / lx + ph[4]) * np.cos(tp * Y / ly + ph[5]) * np.sin(tp * Z / lz)
    return vel


def set_bc_ghosts(vel, n, bclo, bchi, inflow=None):
    """vel.setBndry(0) + inflow fill (:137-163): ghost layers of non-periodic faces are zero except the inflow faces"""
    for d in range(3):
        ax = 3 - d
        for side, bc in ((0, bclo[d]), (1, bchi[d])):
            if bc == PER:
                continue
            sl = [slice(None)] * 4
            sl[ax] = 0 if side == 0 else n[d] + 1
            vel[tuple(sl)] = 0.0
            if bc == INF and inflow is not None:
                sl[0] = d
                vel[tuple(sl)] = inflow(d, side)
    return vel


def save(name, n, dx, bclo, bchi, geom, vel, sigma, out, eb_vel=None):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), n=np.array(n), dx=np.array(dx), bclo=np.array(bclo), bchi=np.array(bchi),
                        vfrac=geom.vfrac, intg=geom.intg, bnorm=geom.bnorm, bintg=geom.bintg, vel=vel,
                        sigma=np.asarray(sigma, dtype=np.float64), eb_vel=np.zeros(0) if eb_vel is None else eb_vel,
                        phi=out["phi"], rhs=out["rhs"], gphi=out["gphi"], vel_new=out["vel_new"])
    cut = geom.cut_mask()
    print(f"{name}: n={n} cut cells={cut.sum()} covered={np.sum(geom.vfrac == 0)} min vfrac={geom.vfrac[cut].min():.3e} "
          f"|vfrac - quadrature| = {np.abs(out['vfrac_quad'] - geom.vfrac).max():.2e} direct-solve residual={out['resid']:.2e} "
          f"|phi|max={np.abs(out['phi']).max():.4f}")


def main():
    rng = np.random.default_rng(20261017)
    # 1. channel_cylinder-x in small: mass inflow x-lo (parabolic profile, probtype 31), pressure outflow x-hi, walls y, periodic z
    n, h = (48, 16, 8), 0.025
    geom = eg.cylinder(n, h, 0.1000001, (0.351, 0.2, 0.0), direction=2)
    bclo, bchi = (INF, NEU, PER), (DIR, NEU, PER)
    vel = smooth_velocity(n, h, rng)
    y = (np.arange(-1, n[1] + 1) + 0.5) * h / (n[1] * h)
    prof = (6.0 * y * (1.0 - y))[None, :] * np.ones((n[2] + 2, 1))
    set_bc_ghosts(vel, n, bclo, bchi, lambda d, side: prof)
    vel[:, 1:-1, 1:-1, 1:-1] *= (geom.vfrac > 0)
    save("eb_channel_cylinder", n, (h,) * 3, bclo, bchi, geom, vel, 1.0, project(n, (h,) * 3, bclo, bchi, geom, vel, 1.0))
    # 2. sphere in a triply periodic box, variable sigma (singular)
    n, h = (16, 16, 16), 1.0 / 16
    geom = eg.sphere(n, h, 0.2300001, (0.47, 0.52, 0.55), small_vfrac=5e-3)
    bclo = bchi = (PER, PER, PER)
    vel = smooth_velocity(n, h, rng, base=(0.5, 0.2, -0.1))
    vel[:, 1:-1, 1:-1, 1:-1] *= (geom.vfrac > 0)
    sigma = rng.uniform(1.0, 4.0, size=geom.vfrac.shape)
    save("eb_sphere_periodic_var", n, (h,) * 3, bclo, bchi, geom, vel, sigma, project(n, (h,) * 3, bclo, bchi, geom, vel, sigma))
    # 3. inclined plane (a ramp) in a closed box: all walls (singular), variable sigma
    n, h = (16, 12, 8), 1.0 / 16
    geom = eg.plane(n, h, (0.0, 0.27, 0.0), (0.35, -1.0, 0.2), small_vfrac=5e-3)
    bclo = bchi = (NEU, NEU, NEU)
    vel = smooth_velocity(n, h, rng, base=(0.0, 0.0, 0.0))
    set_bc_ghosts(vel, n, bclo, bchi)
    vel[:, 1:-1, 1:-1, 1:-1] *= (geom.vfrac > 0)
    sigma = rng.uniform(1.0, 4.0, size=geom.vfrac.shape)
    save("eb_ramp_walls_var", n, (h,) * 3, bclo, bchi, geom, vel, sigma, project(n, (h,) * 3, bclo, bchi, geom, vel, sigma))
    # 4. EB inflow (eb_flow.vel_mag through the cylinder surface), outflow x-hi, everything else walls / periodic
    n, h = (32, 16, 4), 1.0 / 16
    geom = eg.cylinder(n, h, 0.2000001, (0.8, 0.49, 0.0), direction=2)
    bclo, bchi = (NEU, NEU, PER), (DIR, NEU, PER)
    vel = np.zeros((3, n[2] + 2, n[1] + 2, n[0] + 2))
    eb_vel = -0.7 * geom.bnorm * geom.cut_mask()                 # set_eb_velocity with eb_flow.vel_mag = 0.7 (:263-272)
    save("eb_cylinder_ebflow", n, (h,) * 3, bclo, bchi, geom, vel, 0.5, project(n, (h,) * 3, bclo, bchi, geom, vel, 0.5, eb_vel), eb_vel)


if __name__ == "__main__":
    main()
