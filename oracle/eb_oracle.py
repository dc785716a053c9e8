"""CPU oracle (numpy) of the EB nodal projection: Hydro::NodalProjector over amrex::MLMG / MLNodeLaplacian built WITH an
EBFArrayBoxFactory, as incflo drives it when AMREX_USE_EB is on (src/projection/incflo_apply_nodal_projection.cpp:130-136,
:181-201, :215-219; BASELINE configs[4] test_3d/benchmark.channel_cylinder-x).

TEST INFRASTRUCTURE ONLY -- nothing under incflo_b200/ imports this module.

PARITY UNPINNED: MLNodeLaplacian's EB branch (buildIntegral, buildStencil / mlndlap_set_connection / mlndlap_set_stencil_eb,
mlndlap_divu_eb, mlndlap_mknewu_eb, the *_sten smoother / operator and the *_rap transfer operators) lives in AMReX, which is
not vendored with the reference and not present in this image; the reference holds no golden vectors for it.  What this file
restates is the DISCRETISATION those kernels implement, derived from its definition rather than from their text:

  unknowns   phi at nodes, trilinear shape functions N_a on each cell; F_c = fluid part of cell c, given by the volume fraction
             and 18 monomial integrals over F_c in cell-local coordinates (MLNodeLaplacian::m_integral; order in
             incflo_b200/eb_geometry.py);
  operator   (L phi)_a = - sum_c sigma_c sum_b [ int_{F_c} grad N_a . grad N_b ] phi_b   (per unit cell volume): the Q1
             stiffness matrix integrated over the fluid only.  Every integral is a polynomial in x, y, z of degree <= 2 per
             direction, i.e. a combination of the 18 integrals -- AMReX's 27 "connection" coefficients per cell are exactly these
             combinations.  For an uncut cell this is SURVEY A.3's 27-point operator;
  rhs        (D u)_a = - sum_c u_c . int_{F_c} grad N_a   (mlndlap_divu_eb), + EB inflow: sum_c (u_eb . n)_c int_{EB_c} N_a dA
             (setEBInflowVelocity, :196-201; n = boundary normal, out of the fluid);
  update     u_c -= sigma_c * (1/V_c) int_{F_c} grad phi, grad phi = that cell average (mlndlap_mknewu_eb); covered cells: 0;
  rows       NATURAL finite-element rows: at Neumann / inflow faces neither the operator nor the rhs is doubled (AMReX doubles
             both for the "Sigma" strategy and undoes it for RAP -- unimposeNeumannBC).  Inflow faces enter through the normal
             velocity of the first ghost layer exactly as in SURVEY A.2;
  multigrid  the "RAP" coarsening strategy MLNodeLaplacian switches to with EB: Galerkin coarse operators A_c = R A P with
             P = trilinear interpolation, R = P^T / 8 (full weighting), stored as symmetric 27-point stencils (13 forward entries
             + diagonal per node); smoother = Gauss-Seidel in 8 colours c = (i&1) + 2(j&1) + 4(k&1) (mlndlap_gscolor_sten),
             smooth_num_sweeps sweeps per smooth call; nodes with a zero diagonal (covered by the body, or Dirichlet) hold 0;
             MLMG V(nu1, nu2) cycle, BiCGStab bottom solve, solvability offset over the active nodes when no face is Dirichlet
             (SURVEY A.9, A.10).  AMReX's RAP uses stencil-weighted transfer operators; with trilinear P the hierarchy here is a
             different (equally Galerkin) one, so V-cycle counts are this implementation's, the converged answer is not affected.

It is pinned by tests/golden/eb_*.npz: the same discrete problem assembled by integrating the shape-function products
directly over the cut-cell polyhedra (no monomial integrals) and solved with a sparse direct solver
(tests/golden/make_golden_eb.py), and by algebraic identities (tests/test_eb_oracle.py).
Arrays: cells (nz, ny, nx); nodes (nnz, nny, nnx) with nn = n in a periodic direction (unique nodes), n + 1 otherwise.
"""
import itertools

import numpy as np

PER, NEU, DIR, INF = 0, 1, 2, 3          # enum b200np_bc
AX = (2, 1, 0)                           # direction -> numpy axis

# exponents of the 18 volume integrals (amrex i_S_*), see incflo_b200/eb_geometry.py
S_EXP = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (2, 0, 0), (0, 2, 0), (0, 0, 2), (1, 1, 0), (1, 0, 1), (0, 1, 1),
         (2, 1, 0), (2, 0, 1), (1, 2, 0), (0, 2, 1), (1, 0, 2), (0, 1, 2), (2, 2, 0), (2, 0, 2), (0, 2, 2)]
S_IDX = {e: m for m, e in enumerate(S_EXP)}

# the 27 offsets (di, dj, dk), t = (di+1) + 3 (dj+1) + 9 (dk+1); forward ones t = 14..26 are stored (slot m = t - 14),
# slot 13 = diagonal; the backward entry A(p, p - off) is the forward entry of node p - off (symmetry)
OFFS = [(t % 3 - 1, (t // 3) % 3 - 1, t // 9 - 1) for t in range(27)]
FWD = OFFS[14:]
NST = 14
CORNERS = [(a & 1, (a >> 1) & 1, (a >> 2) & 1) for a in range(8)]   # local node a = ax + 2 ay + 4 az


class Params:
    def __init__(self, n, dx, bclo, bchi, max_coarsening_level=100, maxiter=100, nu1=2, nu2=2, nsweeps=4, bottom_maxiter=100,
                 bottom_rtol=1e-4, bottom_atol=-1.0, verbose=0):
        self.n, self.dx = tuple(int(x) for x in n), tuple(float(x) for x in dx)
        self.bclo, self.bchi = tuple(int(x) for x in bclo), tuple(int(x) for x in bchi)
        self.max_coarsening_level, self.maxiter, self.nu1, self.nu2, self.nsweeps = max_coarsening_level, maxiter, nu1, nu2, nsweeps
        self.bottom_maxiter, self.bottom_rtol, self.bottom_atol, self.verbose = bottom_maxiter, bottom_rtol, bottom_atol, verbose
        for d in range(3):
            assert (self.bclo[d] == PER) == (self.bchi[d] == PER)
        self.per = tuple(b == PER for b in self.bclo)
        self.singular = all(b != DIR for b in self.bclo + self.bchi)


def moment(V, S, e):
    return V if e == (0, 0, 0) else S[S_IDX[e]]


def _expo(d, p, q):
    """exponent triple with power p, q in the two directions other than d"""
    e = [0, 0, 0]
    o = [x for x in range(3) if x != d]
    e[o[0]], e[o[1]] = p, q
    return tuple(e)


def grad_integrals(V, S):
    """G[d][a] = int_F d_d N_a  (cell arrays), N_a = prod (1/2 + s x), s = +-1"""
    G = [[None] * 8 for _ in range(3)]
    for d in range(3):
        o = [x for x in range(3) if x != d]
        for a, c in enumerate(CORNERS):
            s = [2 * x - 1 for x in c]
            acc = 0.0
            for p, cp in enumerate((0.5, s[o[0]])):
                for q, cq in enumerate((0.5, s[o[1]])):
                    acc = acc + cp * cq * moment(V, S, _expo(d, p, q))
            G[d][a] = s[d] * acc
    return G


def element_entries(V, S, sigma, dxinv):
    """E[a][b] = - sigma * sum_d dxinv_d^2 int_F d_d N_a d_d N_b  (cell arrays): the contribution of a cell to L(a, b)"""
    E = [[None] * 8 for _ in range(8)]
    for a, ca in enumerate(CORNERS):
        sa = [2 * x - 1 for x in ca]
        for b, cb in enumerate(CORNERS):
            if b < a:
                E[a][b] = E[b][a]
                continue
            sb = [2 * x - 1 for x in cb]
            tot = 0.0
            for d in range(3):
                o = [x for x in range(3) if x != d]
                ce = (0.25, 0.5 * (sa[o[0]] + sb[o[0]]), sa[o[0]] * sb[o[0]])
                cf = (0.25, 0.5 * (sa[o[1]] + sb[o[1]]), sa[o[1]] * sb[o[1]])
                acc = 0.0
                for p in range(3):
                    for q in range(3):
                        if ce[p] != 0 and cf[q] != 0:
                            acc = acc + ce[p] * cf[q] * moment(V, S, _expo(d, p, q))
                tot = tot + dxinv[d] ** 2 * sa[d] * sb[d] * acc
            E[a][b] = -sigma * tot
    return E


def shift(x, off, per):
    """y[p] = x[p + off]: periodic wrap, zero where p + off leaves a non-periodic domain"""
    y = x
    for d in range(3):
        o, ax = off[d], AX[d]
        if o == 0:
            continue
        if per[d]:
            y = np.roll(y, -o, axis=ax)
        else:
            z = np.zeros_like(y)
            src, dst = [slice(None)] * 3, [slice(None)] * 3
            if o > 0:
                src[ax], dst[ax] = slice(o, None), slice(0, -o)
            else:
                src[ax], dst[ax] = slice(0, o), slice(-o, None)
            z[tuple(dst)] = y[tuple(src)]
            y = z
    return y


class Level:
    def __init__(self, n, per, dirlo, dirhi):
        self.n = tuple(n)
        self.per = per
        self.nn = tuple(n[d] if per[d] else n[d] + 1 for d in range(3))
        self.shape = self.nn[::-1]
        self.st = np.zeros((NST,) + self.shape)
        dm = np.zeros(self.shape, dtype=bool)                    # Dirichlet nodes
        for d in range(3):
            if per[d]:
                continue
            sl = [slice(None)] * 3
            if dirlo[d]:
                sl[AX[d]] = 0
                dm[tuple(sl)] = True
            if dirhi[d]:
                sl[AX[d]] = self.nn[d] - 1
                dm[tuple(sl)] = True
        self.dmask = dm
        # a colour couples to itself through a periodic wrap over an odd number of nodes: the colour sweep then reads a snapshot
        self.odd_periodic = any(per[d] and self.nn[d] % 2 == 1 for d in range(3))

    def apply_mask(self):
        keep = (~self.dmask).astype(np.float64)
        for m, off in enumerate(FWD):
            self.st[m] *= keep * shift(keep, off, self.per)
        self.st[13] *= keep

    @property
    def active(self):
        return self.st[13] != 0.0

    def full(self, t):
        """A(p, p + OFFS[t]) for all p"""
        if t == 13:
            return self.st[13]
        if t > 13:
            return self.st[t - 14]
        return shift(self.st[26 - t - 14], OFFS[t], self.per)    # forward entry of the node p + off (off backward)

    def matrix(self):
        """the same operator as a scipy CSR matrix (built from the stencil arrays; used for speed only)"""
        if getattr(self, "_A", None) is None:
            import scipy.sparse as sp
            N = int(np.prod(self.shape))
            idx = np.arange(N).reshape(self.shape)
            rows, cols, vals = [idx.ravel()], [idx.ravel()], [self.st[13].ravel()]
            for m, off in enumerate(FWD):
                tgt = idx
                for d in range(3):
                    if off[d]:
                        q = np.arange(self.nn[d]) + off[d]
                        q = q % self.nn[d] if self.per[d] else np.clip(q, 0, self.nn[d] - 1)   # clipped targets carry a zero entry
                        tgt = np.take(tgt, q, axis=AX[d])
                rows += [idx.ravel(), tgt.ravel()]
                cols += [tgt.ravel(), idx.ravel()]
                vals += [self.st[m].ravel(), self.st[m].ravel()]
            self._A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
            kk, jj, ii = np.meshgrid(*[np.arange(m) for m in self.shape], indexing="ij")
            col = ((ii & 1) + 2 * (jj & 1) + 4 * (kk & 1)).ravel()
            self._rows = []
            for c in range(8):
                r = np.nonzero(col == c)[0]
                self._rows.append((r, self._A[r]))
        return self._A

    def apply(self, x):
        return (self.matrix() @ x.ravel()).reshape(self.shape)

    def apply_stencil(self, x):
        y = self.st[13] * x
        for m, off in enumerate(FWD):
            y = y + self.st[m] * shift(x, off, self.per)
            back = tuple(-o for o in off)
            y = y + shift(self.st[m] * x, back, self.per)
        return y


def cells_to_nodes(K, a, lev):
    """node array with K(cell c) at node c + a"""
    out = np.zeros(lev.shape)
    y = K
    sl = [slice(None)] * 3
    for d in range(3):
        if lev.per[d]:
            if a[d]:
                y = np.roll(y, a[d], axis=AX[d])
        else:
            sl[AX[d]] = slice(a[d], a[d] + lev.n[d])
    out[tuple(sl)] = y
    return out


def build_level0(p, sigma, vfrac, intg):
    """mlndlap_set_connection + mlndlap_set_stencil_eb + mlndlap_set_stencil_s0, from the definition"""
    lev = Level(p.n, p.per, [b == DIR for b in p.bclo], [b == DIR for b in p.bchi])
    sig = np.broadcast_to(np.asarray(sigma, dtype=np.float64), vfrac.shape)
    E = element_entries(vfrac, intg, sig, [1.0 / h for h in p.dx])
    for a, ca in enumerate(CORNERS):
        for b, cb in enumerate(CORNERS):
            off = tuple(cb[d] - ca[d] for d in range(3))
            t = (off[0] + 1) + 3 * (off[1] + 1) + 9 * (off[2] + 1)
            if t < 13:
                continue
            lev.st[13 if t == 13 else t - 14] += cells_to_nodes(E[a][b], ca, lev)
    lev.apply_mask()
    return lev


def _w1(t):
    return 1.0 if t == 0 else 0.5


def _sample(x, a, fl, cl):
    """x[2 I + a] on the coarse node set (0 where 2 I + a is outside a non-periodic fine domain)"""
    y = x
    for d in range(3):
        ax = AX[d]
        idx = 2 * np.arange(cl.nn[d]) + a[d]
        if fl.per[d]:
            y = np.take(y, idx % fl.nn[d], axis=ax)
        else:
            ok = (idx >= 0) & (idx < fl.nn[d])
            y = np.take(y, np.clip(idx, 0, fl.nn[d] - 1), axis=ax)
            shp = [1, 1, 1]
            shp[ax] = -1
            y = y * ok.reshape(shp)
    return y


def coarsen(fl):
    """Galerkin coarse level A_c = (1/8) P^T A P, P trilinear"""
    cl = Level(tuple(m // 2 for m in fl.n), fl.per, [False] * 3, [False] * 3)
    cl.dmask = _sample(fl.dmask.astype(np.float64), (0, 0, 0), fl, cl) > 0
    acc = np.zeros((27,) + cl.shape)
    full = [fl.full(t) for t in range(27)]
    for a in itertools.product((-1, 0, 1), repeat=3):
        wa = _w1(a[0]) * _w1(a[1]) * _w1(a[2])
        for t, o in enumerate(OFFS):
            A_fo = _sample(full[t], a, fl, cl) * (wa * 0.125)
            g = [a[d] + o[d] for d in range(3)]                  # g - 2 I
            # coarse targets J = I + D with |g - 2 D| <= 1 per direction
            opts = []
            for d in range(3):
                opts.append([D for D in (-1, 0, 1) if abs(g[d] - 2 * D) <= 1])
            for D in itertools.product(*opts):
                w = _w1(g[0] - 2 * D[0]) * _w1(g[1] - 2 * D[1]) * _w1(g[2] - 2 * D[2])
                tD = (D[0] + 1) + 3 * (D[1] + 1) + 9 * (D[2] + 1)
                acc[tD] += A_fo * w
    cl.st[13] = acc[13]
    for m in range(13):
        cl.st[m] = acc[14 + m]
    # entries towards nodes outside a non-periodic coarse domain do not exist
    ones = np.ones(cl.shape)
    for m, off in enumerate(FWD):
        cl.st[m] *= shift(ones, off, cl.per)
    cl.apply_mask()
    return cl


def gs_sweeps(lev, x, rhs, nsweeps):
    """mlndlap_gscolor_sten: 8 colours, in place; x = 0 where the diagonal vanishes"""
    lev.matrix()
    d = lev.st[13].ravel()
    act = d != 0.0
    dinv = np.where(act, 1.0 / np.where(act, d, 1.0), 0.0)
    x = x.ravel().copy()
    b = rhs.ravel()
    for _ in range(nsweeps):
        for r, Ar in lev._rows:
            if r.size:                                           # all nodes of a colour at once (= snapshot semantics where a colour
                x[r] = np.where(act[r], x[r] + (b[r] - Ar @ x) * dinv[r], 0.0)   # couples to itself: odd periodic extents)
    return x.reshape(lev.shape)


def residual(lev, x, rhs):
    return np.where(lev.active, rhs - lev.apply(x), 0.0)


def restrict(fl, cl, r):
    out = np.zeros(cl.shape)
    for a in itertools.product((-1, 0, 1), repeat=3):
        out += _sample(r, a, fl, cl) * (_w1(a[0]) * _w1(a[1]) * _w1(a[2]) * 0.125)
    return np.where(cl.active, out, 0.0)


def interp_add(fl, cl, xf, xc):
    """xf += P xc on the active fine nodes"""
    y = xc
    for d in range(3):
        ax = AX[d]
        nf, nc = fl.nn[d], cl.nn[d]
        idx = np.arange(nf)
        lo = idx // 2
        hi = (idx + 1) // 2
        if fl.per[d]:
            hi = hi % nc
        y = 0.5 * (np.take(y, lo, axis=ax) + np.take(y, hi, axis=ax))
    return np.where(fl.active, xf + y, xf)


def bicgstab(lev, b, maxiter, rtol, atol):
    """MLCGSolver::solve_bicgstab, x0 = 0, plain dot products over the nodes; returns (x, iters, code)"""
    x = np.zeros_like(b)
    r = b.copy()
    rh = r.copy()
    rnorm0 = np.abs(r).max()
    if rnorm0 == 0.0 or rnorm0 < atol:
        return x, 0, 0
    rho1 = alpha = omega = 0.0
    p = v = None
    ret, it = 0, 0
    rnorm = rnorm0
    for it in range(1, maxiter + 1):
        rho = np.vdot(rh, r)
        if rho == 0.0:
            ret = 1
            break
        if it == 1:
            p = r.copy()
        else:
            beta = (rho / rho1) * (alpha / omega)
            p = r + beta * (p - omega * v)
        v = lev.apply(p)
        rhtv = np.vdot(rh, v)
        if rhtv == 0.0:
            ret = 2
            break
        alpha = rho / rhtv
        x = x + alpha * p
        s = r - alpha * v
        rnorm = np.abs(s).max()
        if rnorm < rtol * rnorm0 or rnorm < atol:
            break
        t = lev.apply(s)
        t2 = np.vdot(t, t)
        if t2 == 0.0:
            ret = 3
            break
        omega = np.vdot(t, s) / t2
        x = x + omega * s
        r = s - omega * t
        rnorm = np.abs(r).max()
        if rnorm < rtol * rnorm0 or rnorm < atol:
            break
        if omega == 0.0:
            ret = 4
            break
        rho1 = rho
    if ret == 0 and not (rnorm < rtol * rnorm0 or rnorm < atol):
        ret = 8
    return x, it, ret


class MG:
    def __init__(self, p, sigma, vfrac, intg):
        self.p = p
        self.lv = [build_level0(p, sigma, vfrac, intg)]
        while len(self.lv) - 1 < p.max_coarsening_level and all(m % 2 == 0 and m // 2 >= 2 for m in self.lv[-1].n):
            self.lv.append(coarsen(self.lv[-1]))
        self.bottom_iters = 0

    def sub_mean(self, lev, r):
        act = lev.active
        return np.where(act, r - r[act].sum() / act.sum(), 0.0)

    def bottom(self, rhs):
        lev, p = self.lv[-1], self.p
        b = self.sub_mean(lev, rhs) if p.singular else rhs
        x, it, ret = bicgstab(lev, b, p.bottom_maxiter, p.bottom_rtol, p.bottom_atol)
        self.bottom_iters += it
        if ret != 0:                                             # MLMG::bottomSolve: start over with 8 smooth calls
            x = np.zeros_like(b)
            for _ in range(8):
                x = gs_sweeps(lev, x, b, p.nsweeps)
        return x

    def vcycle(self, res0):
        p, nl = self.p, len(self.lv)
        res, cor = [None] * nl, [None] * nl
        res[0] = res0
        for l in range(nl - 1):
            L = self.lv[l]
            cor[l] = np.zeros(L.shape)
            for _ in range(p.nu1):
                cor[l] = gs_sweeps(L, cor[l], res[l], p.nsweeps)
            res[l + 1] = restrict(L, self.lv[l + 1], residual(L, cor[l], res[l]))
        cor[nl - 1] = self.bottom(res[nl - 1])
        for l in range(nl - 2, -1, -1):
            L = self.lv[l]
            cor[l] = interp_add(L, self.lv[l + 1], cor[l], cor[l + 1])
            for _ in range(p.nu2):
                cor[l] = gs_sweeps(L, cor[l], res[l], p.nsweeps)
        return cor[0]

    def solve(self, rhs, rtol, atol):
        """MLMG::solve with a zero initial guess; returns (phi, info)"""
        p, L0 = self.p, self.lv[0]
        rhs = np.where(L0.active, rhs, 0.0)
        if p.singular:
            rhs = self.sub_mean(L0, rhs)
        sol = np.zeros(L0.shape)
        rhsnorm = np.abs(rhs).max()
        res = residual(L0, sol, rhs)
        resnorm0 = np.abs(res).max()
        maxnorm = max(rhsnorm, resnorm0)
        target = max(atol, max(rtol, 1e-16) * maxnorm)
        info = dict(iters=0, rhsnorm=rhsnorm, resnorm0=resnorm0, resnorm=resnorm0, hist=[resnorm0], nlevels=len(self.lv), rhs=rhs)
        if resnorm0 <= target:
            return sol, info
        for it in range(p.maxiter):
            if len(self.lv) == 1:
                cor = self.bottom(res)
            else:
                cor = self.vcycle(res)
            sol = sol + cor
            res = residual(L0, sol, rhs)
            rn = np.abs(res).max()
            info["iters"], info["resnorm"] = it + 1, rn
            info["hist"].append(rn)
            if p.verbose >= 2:
                print(f"MLMG: Iteration {it + 1:3d} Fine resid/bnorm = {rn / maxnorm:.12g}")
            if rn <= target:
                break
            if not rn <= 1e20 * maxnorm:
                raise RuntimeError("MLMG failing so lets stop here")
        else:
            raise RuntimeError("MLMG failed to converge")
        info["bottom_iters"] = self.bottom_iters
        return sol, info


def extend_geom(p, x):
    """cell array + 1 ghost layer: periodic wrap / copy of the adjacent interior cell"""
    y = x
    for d in range(3):
        y = np.concatenate([np.take(y, [-1], axis=AX[d]), y, np.take(y, [0], axis=AX[d])], axis=AX[d]) if p.per[d] else \
            np.concatenate([np.take(y, [0], axis=AX[d]), y, np.take(y, [-1], axis=AX[d])], axis=AX[d])
    return y


def prepare_velocity(p, vel):
    """vel: (3, nz+2, ny+2, nx+2), one ghost layer.  Periodic ghosts are filled (FillBoundary); beyond ONE non-periodic face only
    the normal component survives (the divergence must not see tangential ghost velocity), beyond two or three faces
    nothing does."""
    u = np.array(vel, dtype=np.float64)
    n = p.n
    for d in range(3):
        if p.per[d]:
            ax = AX[d] + 1
            sl_g0, sl_gn, sl_0, sl_n = ([slice(None)] * 4 for _ in range(4))
            sl_g0[ax], sl_gn[ax], sl_0[ax], sl_n[ax] = 0, n[d] + 1, 1, n[d]
            u[tuple(sl_g0)] = u[tuple(sl_n)]
            u[tuple(sl_gn)] = u[tuple(sl_0)]
    kk, jj, ii = np.meshgrid(np.arange(n[2] + 2), np.arange(n[1] + 2), np.arange(n[0] + 2), indexing="ij")
    out = [((ii == 0) | (ii == n[0] + 1)) & (not p.per[0]), ((jj == 0) | (jj == n[1] + 1)) & (not p.per[1]),
           ((kk == 0) | (kk == n[2] + 1)) & (not p.per[2])]
    nout = out[0].astype(int) + out[1].astype(int) + out[2].astype(int)
    for d in range(3):
        keep = (nout == 0) | ((nout == 1) & out[d])
        u[d] = np.where(keep, u[d], 0.0)
    return u


def _ext_nodes(p, K, a, lev):
    """scatter an extended cell array (1 ghost layer) to the nodes: cell c (extended index c + 1) -> node c + a"""
    out = np.zeros(lev.shape)
    for d in range(3):
        assert K.shape[AX[d]] == p.n[d] + 2
    # cells c in [-1, n] ; node index c + a must lie in [0, nn) (non-periodic) -- periodic ghosts duplicate interior cells: skip them
    sl_src, sl_dst = [None] * 3, [None] * 3
    y = K
    for d in range(3):
        ax = AX[d]
        if p.per[d]:
            y = np.take(y, np.arange(1, p.n[d] + 1), axis=ax)   # interior cells only
            if a[d]:
                y = np.roll(y, a[d], axis=ax)
            sl_dst[ax] = slice(None)
        else:
            # cells c = -1 .. n -> nodes c + a; valid nodes 0 .. n
            c_lo = max(-1, -a[d])
            c_hi = min(p.n[d], p.n[d] - a[d])
            y = np.take(y, np.arange(c_lo + 1, c_hi + 2), axis=ax)
            sl_dst[ax] = slice(c_lo + a[d], c_hi + a[d] + 1)
    out[tuple(sl_dst)] = y
    return out


def compute_rhs(p, lev, vel, vfrac, intg, eb_vn=None, bintg=None):
    """rhs = D u (+ EB inflow), natural rows; 0 on inactive nodes"""
    u = prepare_velocity(p, vel)
    Ve = extend_geom(p, vfrac)
    Se = np.stack([extend_geom(p, intg[m]) for m in range(18)])
    G = grad_integrals(Ve, Se)
    dxinv = [1.0 / h for h in p.dx]
    rhs = np.zeros(lev.shape)
    for a, ca in enumerate(CORNERS):
        contrib = 0.0
        for d in range(3):
            contrib = contrib - dxinv[d] * u[d] * G[d][a]
        rhs += _ext_nodes(p, contrib, ca, lev)
    if eb_vn is not None:
        B = bintg
        for a, ca in enumerate(CORNERS):
            s = [2 * x - 1 for x in ca]
            bn = (0.125 * B[0] + 0.25 * (s[0] * B[1] + s[1] * B[2] + s[2] * B[3])
                  + 0.5 * (s[0] * s[1] * B[4] + s[0] * s[2] * B[5] + s[1] * s[2] * B[6]) + s[0] * s[1] * s[2] * B[7])
            rhs += cells_to_nodes(dxinv[0] * eb_vn * bn, ca, lev)
    return np.where(lev.active, rhs, 0.0)


def node_at(phi, a, lev):
    """phi(c + a) as a cell array"""
    y = phi
    for d in range(3):
        ax = AX[d]
        idx = np.arange(lev.n[d]) + a[d]
        if lev.per[d]:
            idx = idx % lev.nn[d]
        y = np.take(y, idx, axis=ax)
    return y


def gradient(p, lev, phi, vfrac, intg):
    """(1/V) int_F grad phi per cell (mlndlap_mknewu_eb / compGrad); 0 in covered cells"""
    G = grad_integrals(vfrac, intg)
    dxinv = [1.0 / h for h in p.dx]
    fluid = vfrac > 0.0
    vinv = np.where(fluid, 1.0 / np.where(fluid, vfrac, 1.0), 0.0)
    g = np.zeros((3,) + vfrac.shape)
    for a, ca in enumerate(CORNERS):
        pa = node_at(phi, ca, lev)
        for d in range(3):
            g[d] += dxinv[d] * pa * G[d][a]
    return g * vinv


def project(p, vel, sigma, vfrac, intg, rtol, atol, eb_vel=None, bnorm=None, bintg=None):
    """NodalProjector::project with an EB factory.  vel (3, nz+2, ny+2, nx+2) with one ghost layer (input at non-periodic
    faces).  Returns dict(phi, gphi, vel (valid cells), info)."""
    mg = MG(p, sigma, vfrac, intg)
    L0 = mg.lv[0]
    eb_vn = None
    if eb_vel is not None:
        eb_vn = eb_vel[0] * bnorm[0] + eb_vel[1] * bnorm[1] + eb_vel[2] * bnorm[2]
    rhs = compute_rhs(p, L0, vel, vfrac, intg, eb_vn, bintg)
    phi, info = mg.solve(rhs, rtol, atol)
    g = gradient(p, L0, phi, vfrac, intg)
    sig = np.broadcast_to(np.asarray(sigma, dtype=np.float64), vfrac.shape)
    u = np.array(vel[:, 1:-1, 1:-1, 1:-1], dtype=np.float64)
    unew = np.where(vfrac > 0.0, u - sig * g, 0.0)
    return dict(phi=phi, gphi=g, vel=unew, info=info, mg=mg, rhs=rhs)
