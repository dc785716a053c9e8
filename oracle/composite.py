"""CPU oracle of the composite (two AMR level) nodal projection -- TEST INFRASTRUCTURE ONLY.

Restates the multi-level branch of the path behind incflo::ApplyNodalProjection
(src/projection/incflo_apply_nodal_projection.cpp:101-121 sigma per level, :137 setBndry(0) on every
level, :181-219 projector over Geom(0,finest_level), :258-266 average_down(gp)) for ONE fine box at
refinement ratio 2 (amr.ref_ratio = 2 in every deck, e.g. test_no_eb_3d/benchmark.bouss_bubble_god).
The arithmetic lives in un-vendored AMReX (MLMG::oneIter multi-level branch, MLNodeLaplacian::reflux /
compRHS / interpolationAmr); "parity unpinned" as for the single-level oracle.
It is pinned by tests/golden/composite_*.npz: an independent SciPy direct solve of the composite Q1
finite-element problem with hanging-node constraints (tests/golden/make_golden_composite.py).

Algorithm (AMReX MLMG with two AMR levels, ref ratio 2 => the fine AMR level has a single MG level, so
its "miniCycle" is nu1 smooth calls):
  unknowns  sol0 on every coarse node, sol1 on every node of the fine box; fine nodes on the box
            boundary are not relaxed (Dirichlet for the fine level) and only change by interpolation of
            the coarse correction, so they always equal the trilinear interpolant of sol0 there.
  rhs       fine interior nodes: D u1.  Coarse nodes outside the box: D u0.  Coarse nodes ON the box
            boundary (coarse/fine interface): the divergence of the uncovered coarse cells plus the
            full-weighting restriction of the divergence of the fine cells inside the box
            (mlndlap_divu_cf_contrib + mlndlap_divu_fine_contrib).  Coarse nodes strictly inside the
            box: restriction of the fine rhs.
  one iteration (MLMG::oneIter):
     cor1 = 0; nu1 smooth calls on the fine level (homogeneous Dirichlet at the interface); sol1 += cor1
     composite coarse residual (computeResWithCrseSolFineCor + reflux); makeSolvable if singular
     coarse V-cycle (single-level machinery) -> cor0; sol0 += cor0
     cor1 = trilinear interpolation of cor0 on every fine node (interpolationAmr); sol1 += cor1
     res1 -= L1 cor1; cor1 = 0; nu2 smooth calls; sol1 += cor1
  convergence: inf-norm of the composite residual (fine interior nodes + coarse nodes not strictly
            inside the box) <= max(atol, rtol * max(rhsnorm, resnorm0)).
  finish    sol0 <- injection of sol1 on covered nodes; u -= sigma G sol and gphi = G sol per level;
            average_down of u1, gphi1 onto the covered coarse cells (:258-266, NodalProjector A.1 (7)).

The one-sided ("fine cells inside the box only") sums needed at the interface are evaluated with the
ordinary single-level kernels on the fine box taken as a domain with reflecting (Neumann) faces: for a
node on f faces of the box the reflected operator / divergence equals 2^f times the one-sided sum
(SURVEY A.2, A.8), and full weighting of the reflected residual gives 2^f times the wanted
restriction.  The uncovered-coarse-cells sums are evaluated with sigma = 0 and u = 0 in the covered
coarse cells.  The CUDA path (incflo_b200/csrc, b200np_composite_*) uses exactly the same construction.
"""
import numpy as np

from . import pyoracle as po

PER, NEU, DIR, INF = 0, 1, 2, 3


def _trilinear(c, periodic=(False, False, False)):
    """coarse nodal values of the box -> fine nodal values (mlmg_lin_nd_interp_r2).  Array axes are (z, y, x);
    periodic[d] (d = x, y, z): the box spans the periodic direction d (n unique nodes -> 2n), else nb+1 -> 2nb+1"""
    f = c
    for ax in range(3):
        per = periodic[2 - ax]
        sh = list(f.shape)
        sh[ax] = 2 * sh[ax] if per else 2 * sh[ax] - 1
        g = np.empty(sh)
        ev = [slice(None)] * 3; od = [slice(None)] * 3; lo = [slice(None)] * 3; hi = [slice(None)] * 3
        ev[ax] = slice(0, None, 2); od[ax] = slice(1, None, 2); lo[ax] = slice(0, -1); hi[ax] = slice(1, None)
        g[tuple(ev)] = f
        g[tuple(od)] = 0.5 * (f + np.roll(f, -1, axis=ax)) if per else 0.5 * (f[tuple(lo)] + f[tuple(hi)])
        f = g
    return f


def _avg_down_cells(f):
    """mean of the 8 children; f has shape (..., 2nz, 2ny, 2nx)"""
    return 0.125 * (f[..., 0::2, 0::2, 0::2] + f[..., 0::2, 0::2, 1::2] + f[..., 0::2, 1::2, 0::2] + f[..., 0::2, 1::2, 1::2] +
                    f[..., 1::2, 0::2, 0::2] + f[..., 1::2, 0::2, 1::2] + f[..., 1::2, 1::2, 0::2] + f[..., 1::2, 1::2, 1::2])


def _face_count(nb):
    """number of box faces each node of the (nb+1)^3 box lies on, shape (nbz+1, nby+1, nbx+1)"""
    cnt = np.zeros((nb[2] + 1, nb[1] + 1, nb[0] + 1), dtype=np.int64)
    cnt[0, :, :] += 1; cnt[-1, :, :] += 1
    cnt[:, 0, :] += 1; cnt[:, -1, :] += 1
    cnt[:, :, 0] += 1; cnt[:, :, -1] += 1
    return cnt


def vcycle(mg, res0, nu1=2, nu2=2, nsweeps=4):
    """MLMG::mgVcycle on (cor, res) from level 0 of a single-level hierarchy; returns cor[0]"""
    nl = mg.nlev
    res = [res0] + [None] * (nl - 1)
    cor = [None] * nl
    for l in range(nl - 1):
        cor[l] = np.zeros(mg.node_shape(l))
        mg.smooth(l, cor[l], res[l], nu1 * nsweeps)
        rescor = mg.residual(l, cor[l], res[l])
        res[l + 1] = mg.restrict(l, rescor)
    cor[nl - 1], _ = mg.bottom_solve(np.ascontiguousarray(res[nl - 1]))
    for l in range(nl - 2, -1, -1):
        mg.interp_add(l, cor[l], cor[l + 1])
        mg.smooth(l, cor[l], res[l], nu2 * nsweeps)
    return cor[0]


class CompositeProjector:
    """Two-level composite nodal projection; coarse domain params0 (pyoracle.Params), fine box =
    coarse cells [clo, chi] (inclusive) refined by 2.  Per direction the box either
      * lies at least one coarse cell inside the domain on that side (a coarse/fine interface), or
      * spans a periodic direction completely (the fine level is periodic there), or
      * touches a Neumann (wall) or Dirichlet (outflow) domain face (the fine level inherits that BC).
    Touching an inflow face or the periodic seam without spanning the direction is not supported."""

    def __init__(self, params0, clo, chi, smoother_kw=None):
        self.p0 = params0
        self.n0 = tuple(params0.n)
        self.dx0 = tuple(params0.dx)
        self.bclo = tuple(params0.bclo); self.bchi = tuple(params0.bchi)
        self.clo = tuple(int(x) for x in clo); self.chi = tuple(int(x) for x in chi)
        n0 = self.n0
        self.span = [False] * 3; self.cf_lo = [True] * 3; self.cf_hi = [True] * 3
        bD_lo, bD_hi, bN_lo, bN_hi = [DIR] * 3, [DIR] * 3, [NEU] * 3, [NEU] * 3
        for d in range(3):
            assert 0 <= self.clo[d] <= self.chi[d] <= n0[d] - 1
            at_lo, at_hi = self.clo[d] == 0, self.chi[d] == n0[d] - 1
            if self.bclo[d] == PER:
                assert at_lo == at_hi, "a fine box may span a periodic direction but not touch its seam"
                if at_lo:
                    self.span[d] = True; self.cf_lo[d] = self.cf_hi[d] = False
                    bD_lo[d] = bD_hi[d] = bN_lo[d] = bN_hi[d] = PER
            else:
                if at_lo:
                    assert self.bclo[d] in (NEU, DIR), "fine box on an inflow face is not supported"
                    self.cf_lo[d] = False; bD_lo[d] = bN_lo[d] = self.bclo[d]
                if at_hi:
                    assert self.bchi[d] in (NEU, DIR), "fine box on an inflow face is not supported"
                    self.cf_hi[d] = False; bD_hi[d] = bN_hi[d] = self.bchi[d]
        assert any(self.cf_lo) or any(self.cf_hi), "the fine box covers the whole domain"
        self.nb = tuple(self.chi[d] - self.clo[d] + 1 for d in range(3))
        self.nf = tuple(2 * x for x in self.nb)
        self.dx1 = tuple(0.5 * x for x in self.dx0)
        kw = dict(smoother_kw or {})
        self.pD = po.make_params(self.nf, self.dx1, tuple(bD_lo), tuple(bD_hi), **kw)   # interface nodes are not relaxed
        self.pN = po.make_params(self.nf, self.dx1, tuple(bN_lo), tuple(bN_hi), **kw)   # reflecting interface: one-sided sums
        self.singular = all(b != DIR for b in self.bclo + self.bchi)
        self.nu1 = self.nu2 = 2
        self.nsweeps = 4
        self.maxiter = 100
        # slices (z, y, x) of the box in the coarse arrays: cells, nodes (unique-node layout), strictly covered nodes
        c, h = self.clo, self.chi
        self.cbox = tuple(slice(c[d], h[d] + 1) for d in (2, 1, 0))
        self.nbox = tuple(slice(0, n0[d]) if self.span[d] else slice(c[d], h[d] + 2) for d in (2, 1, 0))
        self.nint = tuple(slice(c[d] + (1 if self.cf_lo[d] else 0), (h[d] + 1) if self.cf_hi[d] else (n0[d] if self.span[d] else h[d] + 2))
                          for d in (2, 1, 0))
        # number of coarse/fine interface faces each box node lies on -> 1 / 2^f; Dirichlet mask of the fine level
        shape = tuple(n0[d] if self.span[d] else self.nb[d] + 1 for d in (2, 1, 0))
        fc = np.zeros(shape, dtype=np.int64)
        fshape = tuple(self.nf[d] if self.span[d] else self.nf[d] + 1 for d in (2, 1, 0))
        self.fmask = np.zeros(fshape, dtype=bool)      # fine nodes that are not unknowns of the fine level
        for ax, d in enumerate((2, 1, 0)):
            lo = [slice(None)] * 3; hi = [slice(None)] * 3
            lo[ax] = 0; hi[ax] = -1
            if self.cf_lo[d]: fc[tuple(lo)] += 1
            if self.cf_hi[d]: fc[tuple(hi)] += 1
            if bD_lo[d] == DIR: self.fmask[tuple(lo)] = True
            if bD_hi[d] == DIR: self.fmask[tuple(hi)] = True
        self.scale = 0.5 ** fc

    # ---- composite pieces --------------------------------------------------------------------
    def _add_fine_part(self, r0, RN):
        """coarse array r0 (uncovered-cell sums) += restricted fine-side sums RN (box nodes)"""
        r0[self.nbox] += self.scale * RN
        return r0

    def _coarse_residual(self, sol0, sol1):
        """composite residual on the coarse level (reflux): rhs0 - A_composite(sol0, sol1)"""
        rN = self.mgN.residual(0, sol1, self.rhsN)            # 2^f x one-sided (b - A) on the box boundary, full inside
        RN = self.mgN.restrict(0, rN)
        r0 = self.mg0z.residual(0, sol0, self.rhs0z)          # uncovered coarse cells only
        r0 = self._add_fine_part(r0, RN)
        return r0 - self.offset

    def _norm(self, res0, res1):
        m = np.ones(res0.shape, dtype=bool)
        m[self.nint] = False                                 # strictly covered coarse nodes do not count
        return max(np.abs(res0[m]).max(), np.abs(res1).max())

    # ---- Hydro::NodalProjector::project over two levels -------------------------------------------
    def setup(self, sigma0=None, sigma1=None, const_sigma=1.0):
        """the four single-level hierarchies: coarse (sigma averaged down under the box), coarse with sigma = 0 in
        the covered cells, fine box with Dirichlet / with reflecting interface faces"""
        n0, nf = self.n0, self.nf
        s1 = np.full((nf[2], nf[1], nf[0]), float(const_sigma)) if sigma1 is None else np.ascontiguousarray(sigma1, dtype=np.float64)
        s0 = np.full((n0[2], n0[1], n0[0]), float(const_sigma)) if sigma0 is None else np.array(sigma0, dtype=np.float64)
        s0[self.cbox] = _avg_down_cells(s1)                  # averageDownCoeffsToCoarseAmrLevel
        s0z = s0.copy(); s0z[self.cbox] = 0.0
        var = sigma0 is not None or sigma1 is not None
        self.mg0 = po.MG(self.p0, s0 if var else None, const_sigma)
        self.mg0z = po.MG(self.p0, s0z, 1.0)
        self.mgD = po.MG(self.pD, s1 if var else None, const_sigma)
        self.mgN = po.MG(self.pN, s1 if var else None, const_sigma)
        self.rhsN = np.zeros(self.mgN.node_shape(0)); self.rhs0z = np.zeros(self.mg0.node_shape(0)); self.offset = 0.0

    def fill_hanging(self, sol0, sol1):
        """fine nodes on the coarse/fine interface <- trilinear interpolant of the coarse solution (what the
        iteration maintains implicitly: they only ever change by interpolationAmr)"""
        interp = _trilinear(sol0[self.nbox], self.span)
        hang = np.zeros(sol1.shape, dtype=bool)
        for ax, d in enumerate((2, 1, 0)):
            lo = [slice(None)] * 3; hi = [slice(None)] * 3
            lo[ax] = 0; hi[ax] = -1
            if self.cf_lo[d]: hang[tuple(lo)] = True
            if self.cf_hi[d]: hang[tuple(hi)] = True
        sol1[hang] = interp[hang]
        return hang

    def composite_residual(self, sol0, sol1):
        """(res0, res1) = rhs - A_composite(sol0, sol1) with the current rhs (zero after setup())"""
        rhs1 = getattr(self, "rhs1_off", None)
        if rhs1 is None:
            rhs1 = np.zeros_like(sol1)
        return self._coarse_residual(sol0, sol1), self.mgD.residual(0, sol1, rhs1)

    def project(self, vel0, ng0, vel1, ng1, sigma0=None, sigma1=None, const_sigma=1.0, rtol=1e-11, atol=1e-14):
        """vel0 (3, n0z+2ng0, ...) and vel1 (3, nfz+2ng1, ...) are updated in place (valid cells).
        sigma0 / sigma1: cell arrays or None => const_sigma on both levels."""
        n0, nf, nb = self.n0, self.nf, self.nb
        self.setup(sigma0, sigma1, const_sigma)
        # velocities: ghost cells of the fine level are zero (vel.setBndry(0.0), :137), covered coarse cells do not count
        v1 = vel1.copy()
        gz = np.zeros_like(v1)
        gz[:, ng1:ng1 + nf[2], ng1:ng1 + nf[1], ng1:ng1 + nf[0]] = v1[:, ng1:ng1 + nf[2], ng1:ng1 + nf[1], ng1:ng1 + nf[0]]
        v0z = vel0.copy()
        v0z[:, ng0 + self.clo[2]:ng0 + self.chi[2] + 1, ng0 + self.clo[1]:ng0 + self.chi[1] + 1,
            ng0 + self.clo[0]:ng0 + self.chi[0] + 1] = 0.0
        self.rhs1 = self.mgD.divu(gz, ng1)                    # 0 on the box boundary (Dirichlet nodes)
        self.rhsN = self.mgN.divu(gz, ng1)                    # 2^f x one-sided divergence on the box boundary
        self.rhs0z = self.mg0.divu(v0z, ng0)
        # composite coarse rhs (only used for the solvability offset and the norms)
        rhs0 = self._add_fine_part(self.rhs0z.copy(), self.mgN.restrict(0, self.rhsN))
        self.offset = 0.0
        if self.singular:                                    # MLMG::makeSolvable: one offset for every level
            w = self.mg0.dot_weights(0)
            self.offset = float((w * rhs0).sum() / w.sum())
        rhs0 -= self.offset
        rhs1 = self.rhs1.copy()
        rhs1[~self.fmask] -= self.offset
        sol0 = np.zeros_like(rhs0)
        sol1 = np.zeros_like(rhs1)
        res1 = self.mgD.residual(0, sol1, rhs1)
        res0 = self._coarse_residual(sol0, sol1)
        m = np.ones(rhs0.shape, dtype=bool); m[self.nint] = False
        rhsnorm = max(np.abs(rhs0[m]).max(), np.abs(rhs1).max())
        resnorm0 = self._norm(res0, res1)
        target = max(atol, max(rtol, 1e-16) * max(rhsnorm, resnorm0))
        hist = [resnorm0]
        iters, status = 0, 0
        if resnorm0 > target:
            status = 1
            for it in range(self.maxiter):
                # fine level: miniCycle = nu1 smooth calls on (cor, res)
                cor1 = np.zeros_like(sol1)
                self.mgD.smooth(0, cor1, res1, self.nu1 * self.nsweeps)
                sol1 += cor1
                # coarse level: composite residual, solvability, V-cycle
                res0 = self._coarse_residual(sol0, sol1)
                if self.singular:
                    w = self.mg0.dot_weights(0) if it == 0 else w
                    res0 -= (w * res0).sum() / w.sum()
                cor0 = vcycle(self.mg0, np.ascontiguousarray(res0), self.nu1, self.nu2, self.nsweeps)
                sol0 += cor0
                # interpolate the coarse correction to every fine node, then post-smooth its residual
                cor1 = np.ascontiguousarray(_trilinear(cor0[self.nbox], self.span))
                sol1 += cor1
                res1 = self.mgD.residual(0, sol1, rhs1)
                cor1 = np.zeros_like(sol1)
                self.mgD.smooth(0, cor1, res1, self.nu2 * self.nsweeps)
                sol1 += cor1
                # convergence on the composite residual
                res1 = self.mgD.residual(0, sol1, rhs1)
                res0 = self._coarse_residual(sol0, sol1)
                rn = self._norm(res0, res1)
                hist.append(rn)
                iters = it + 1
                if rn <= target:
                    status = 0
                    break
                if not rn <= 1e20 * max(rhsnorm, resnorm0):
                    status = 2
                    break
        # finish: injection, velocity / gradient update, average down
        sol0[self.nbox] = sol1[::2, ::2, ::2]
        g1 = self.mgD.mknewu(sol1, vel1, ng1)                 # vel1 -= sigma1 G sol1 ; g1 = G sol1
        g0 = self.mg0.mknewu(sol0, vel0, ng0)
        g0[(slice(None),) + self.cbox] = _avg_down_cells(g1)
        vel0[:, ng0 + self.clo[2]:ng0 + self.chi[2] + 1, ng0 + self.clo[1]:ng0 + self.chi[1] + 1,
             ng0 + self.clo[0]:ng0 + self.chi[0] + 1] = _avg_down_cells(
            vel1[:, ng1:ng1 + nf[2], ng1:ng1 + nf[1], ng1:ng1 + nf[0]])
        return dict(status=status, iters=iters, phi0=sol0, phi1=sol1, gphi0=g0, gphi1=g1, rhsnorm=rhsnorm,
                    resnorm0=resnorm0, resnorm=hist[-1], hist=hist)


def apply_nodal_projection(cp, velocity, ng, gp, p_nd, density=None, ngd=(0, 0), ro_0=1.0, scaling_factor=1.0,
                           rtol=1e-11, atol=1e-14):
    """incflo::ApplyNodalProjection with finest_level = 1, non-incremental branch (:39-71, :101-121,
    :221-266): per level u += s gp / rho, sigma = s / rho, project, gp = G phi, p_nd = phi (nodal boxes
    [0, n] per level, periodic image planes included), average_down(gp).  velocity / gp / p_nd / density
    are pairs (level 0, level 1), updated in place."""
    n = (cp.n0, cp.nf)
    sig = [None, None]
    for l in range(2):
        nx, ny, nz = n[l]
        g = ng[l]
        v = velocity[l]
        inner = (slice(None), slice(g, g + nz), slice(g, g + ny), slice(g, g + nx))
        if density is not None:
            d = ngd[l]
            rho = density[l][d:d + nz, d:d + ny, d:d + nx]
            sig[l] = scaling_factor / rho
            v[inner] += gp[l] * sig[l][None]
        else:
            v[inner] += gp[l] * (scaling_factor / ro_0)
        # vel.setBndry(0.0): ghost cells (no inflow faces in the composite cases)
        keep = v[inner].copy()
        v[...] = 0.0
        v[inner] = keep
    r = cp.project(velocity[0], ng[0], velocity[1], ng[1], sig[0], sig[1], scaling_factor / ro_0, rtol, atol)
    gp[0][...] = r["gphi0"]; gp[1][...] = r["gphi1"]
    p0 = r["phi0"]
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        if cp.bclo[d] == PER:
            p0 = np.concatenate([p0, np.take(p0, [0], axis=ax)], axis=ax)
    p_nd[0][...] = p0
    p_nd[1][...] = r["phi1"]
    return r
