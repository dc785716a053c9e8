"""Golden fixtures tests/golden/mac_*.npz for the MAC projection (Hydro::MacProjector over MLABecLaplacian).

Independent evaluation of the discrete problem, no multigrid and no code shared with oracle/ or incflo_b200/csrc:
the cell-centred 7-point operator -div(b grad phi) is assembled face by face into a scipy.sparse matrix, with the
boundary faces written out explicitly --
  periodic  : the face couples the first and the last cell,
  Neumann   : the face carries no flux,
  Dirichlet : phi = 0 on the face, flux = b (3 phi_0 - phi_1 / 3) / dx  (the quadratic through the face value and the
              first two cell centres, i.e. AMReX's maxorder = 3 ghost cell -2 phi_0 + phi_1 / 3),
rhs = -div(u_mac), sparse direct solve (mean-free rhs and a Lagrange multiplier when no face is Dirichlet), and
u_mac -= b grad phi with the same boundary gradients.
Run: python tests/golden/make_golden_mac.py
"""
import os

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

HERE = os.path.dirname(os.path.abspath(__file__))
PER, NEU, DIR = 0, 1, 2


def assemble(n, dx, bclo, bchi, b):
    nx, ny, nz = n
    N = nx * ny * nz
    idx = np.arange(N).reshape(nz, ny, nx)
    rows, cols, vals = [], [], []

    def add(r, c, v):
        rows.append(np.ravel(r)); cols.append(np.ravel(c)); vals.append(np.ravel(v))
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        h2 = 1.0 / (dx[d] * dx[d])
        m = n[d]
        take = lambda a, s: np.take(a, s, axis=ax)
        # interior faces f = 1 .. m-1 between cells f-1 and f
        lo, hi = take(idx, np.arange(0, m - 1)), take(idx, np.arange(1, m))
        bf = take(b[d], np.arange(1, m)) * h2
        add(lo, lo, bf); add(lo, hi, -bf); add(hi, hi, bf); add(hi, lo, -bf)
        first, last = take(idx, [0]), take(idx, [m - 1])
        b0, bm = take(b[d], [0]) * h2, take(b[d], [m]) * h2
        if bclo[d] == PER:
            add(first, first, b0); add(first, last, -b0); add(last, last, b0); add(last, first, -b0)
        else:
            for bc, cell, nxt, bb in ((bclo[d], first, take(idx, [1]), b0), (bchi[d], last, take(idx, [m - 2]), bm)):
                if bc == DIR:
                    add(cell, cell, 3.0 * bb); add(cell, nxt, -bb / 3.0)
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    return A


def project(n, dx, bclo, bchi, u, v, w, b):
    A = assemble(n, dx, bclo, bchi, b)
    rhs = -((u[:, :, 1:] - u[:, :, :-1]) / dx[0] + (v[:, 1:, :] - v[:, :-1, :]) / dx[1] + (w[1:] - w[:-1]) / dx[2])
    singular = all(x != DIR for x in tuple(bclo) + tuple(bchi))
    r = rhs.ravel().copy()
    if singular:
        r -= r.mean()
        one = sp.csc_matrix(np.ones((r.size, 1)))
        K = sp.bmat([[A, one], [one.T, None]]).tocsc()
        phi = spl.spsolve(K, np.concatenate([r, [0.0]]))[:-1]
    else:
        phi = spl.spsolve(A.tocsc(), r)
    phi = phi.reshape(rhs.shape)
    out = [u.copy(), v.copy(), w.copy()]
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        m = n[d]
        g = np.zeros_like(b[d])
        sl = lambda s: tuple(s if a == ax else slice(None) for a in range(3))
        g[sl(slice(1, m))] = (phi[sl(slice(1, m))] - phi[sl(slice(0, m - 1))]) / dx[d]
        if bclo[d] == PER:
            gw = (phi[sl(0)] - phi[sl(m - 1)]) / dx[d]
            g[sl(0)] = gw; g[sl(m)] = gw
        else:
            if bclo[d] == DIR:
                g[sl(0)] = (3.0 * phi[sl(0)] - phi[sl(1)] / 3.0) / dx[d]
            if bchi[d] == DIR:
                g[sl(m)] = -(3.0 * phi[sl(m - 1)] - phi[sl(m - 2)] / 3.0) / dx[d]
        out[d] -= b[d] * g
    return dict(phi=phi, u=out[0], v=out[1], w=out[2], rhs=rhs)


def fields(n, seed, var, bclo, bchi):
    rng = np.random.default_rng(seed)
    nx, ny, nz = n

    def smooth(a):
        for ax in range(3):
            a = 0.5 * a + 0.25 * (np.roll(a, 1, ax) + np.roll(a, -1, ax))
        return a
    u = smooth(rng.standard_normal((nz, ny, nx + 1))); v = smooth(rng.standard_normal((nz, ny + 1, nx))); w = smooth(rng.standard_normal((nz + 1, ny, nx)))
    vel = [u, v, w]
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        sl = lambda s: tuple(s if a == ax else slice(None) for a in range(3))
        if bclo[d] == PER:
            vel[d][sl(n[d])] = vel[d][sl(0)]
        else:
            if bclo[d] == NEU: vel[d][sl(0)] = 0.0          # wall (an inflow face would carry the inflow value)
            if bchi[d] == NEU: vel[d][sl(n[d])] = 0.0
    if var:   # dt / rho on faces from a cell density, as average_cellcenter_to_face + invert does (:81-90)
        rho = 1.0 + 1.5 * smooth(rng.uniform(0, 1, size=(nz, ny, nx)))
        dt = 0.05
        b = []
        for d, ax in ((0, 2), (1, 1), (2, 0)):
            per = bclo[d] == PER
            lo = np.take(rho, [n[d] - 1] if per else [0], axis=ax)
            hi = np.take(rho, [0] if per else [n[d] - 1], axis=ax)
            ext = np.concatenate([lo, rho, hi], axis=ax)
            sl = lambda s: tuple(s if a == ax else slice(None) for a in range(3))
            face = 0.5 * (ext[sl(slice(0, -1))] + ext[sl(slice(1, None))])
            b.append(dt / face)
    else:
        b = [np.full(x.shape, 0.37) for x in vel]
    return vel, b


def cases():
    return [
        dict(name="mac_periodic_const", n=(16, 16, 16), dx=(1 / 16,) * 3, bclo=(PER, PER, PER), bchi=(PER, PER, PER), var=False, seed=1),
        dict(name="mac_rt_walls_var", n=(16, 16, 16), dx=(1 / 16,) * 3, bclo=(PER, PER, NEU), bchi=(PER, PER, NEU), var=True, seed=2),
        dict(name="mac_channel_inflow_outflow_var", n=(24, 8, 12), dx=(1 / 24, 1 / 24, 1 / 24), bclo=(NEU, NEU, PER), bchi=(DIR, NEU, PER), var=True, seed=3),
        dict(name="mac_box_aniso_dirichlet_const", n=(12, 8, 10), dx=(0.1, 0.07, 0.05), bclo=(DIR, NEU, NEU), bchi=(DIR, NEU, DIR), var=False, seed=4),
    ]


def main():
    for c in cases():
        (u, v, w), b = fields(c["n"], c["seed"], c["var"], c["bclo"], c["bchi"])
        if c["name"].startswith("mac_channel"):
            u[:, :, 0] = 1.0 + 0.1 * np.random.default_rng(9).standard_normal(u[:, :, 0].shape)   # mass inflow on x-lo
        r = project(c["n"], c["dx"], c["bclo"], c["bchi"], u, v, w, b)
        path = os.path.join(HERE, c["name"] + ".npz")
        np.savez_compressed(path, n=np.array(c["n"]), dx=np.array(c["dx"]), bclo=np.array(c["bclo"]), bchi=np.array(c["bchi"]),
                            u_in=u, v_in=v, w_in=w, bx=b[0], by=b[1], bz=b[2], phi=r["phi"], u_out=r["u"], v_out=r["v"], w_out=r["w"], rhs=r["rhs"])
        div = (r["u"][:, :, 1:] - r["u"][:, :, :-1]) / c["dx"][0] + (r["v"][:, 1:] - r["v"][:, :-1]) / c["dx"][1] + (r["w"][1:] - r["w"][:-1]) / c["dx"][2]
        print(f"{c['name']}: |div u| before {np.abs(r['rhs']).max():.3e} after {np.abs(div - (0 if any(x == DIR for x in c['bclo'] + c['bchi']) else div.mean())).max():.3e}")


if __name__ == "__main__":
    main()
