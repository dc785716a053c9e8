"""CPU oracle (numpy) of the MAC projection: Hydro::MacProjector over amrex::MLMG / MLABecLaplacian as incflo drives
it in src/convection/incflo_compute_MAC_projected_velocities.cpp:69-129 (inv_rho on faces, initProjector /
updateCoeffs / updateBeta, setDomainBC(get_mac_projection_bc)) and :280-299 (project(rtol, atol), project(mac_phi, ..)).

TEST INFRASTRUCTURE ONLY -- nothing under incflo_b200/ imports this module.

PARITY UNPINNED: MacProjector, MLMG, MLABecLaplacian, MLCellLinOp and MLCGSolver live in AMReX / AMReX-Hydro, which are
not vendored with the reference (README.md:14-28) and the reference holds no golden vectors for this path.  This file
restates the published algorithm for ONE AMR level and ONE box:

  operator   A phi = - sum_d (1/dx_d^2) [ b_d(i+1/2) (phi(i+1) - phi(i)) - b_d(i-1/2) (phi(i) - phi(i-1)) ]
             (MLABecLaplacian with alpha = 0, beta = 1, b = dt / rho on faces; mlabeclap_adotx)
  BCs        LinOpBCType of incflo::get_mac_projection_bc (src/projection/incflo_projection_bc.cpp:43-79):
             Periodic; Neumann (walls, mass inflow: ghost = first interior cell); Dirichlet (pressure in/outflow,
             phi = 0 on the face): ghost cell by polynomial extrapolation of order maxorder = 3 through the face
             value and the first two cells, ghost = -2 phi_0 + phi_1 / 3 (mllinop_apply_bc, poly_interp_coeff);
  smoother   red-black Gauss-Seidel with over-relaxation 1.15, the boundary stencil folded into the diagonal
             (abec_gsrb: phi += omega / (gamma - delta) * (rhs - (gamma phi - rho))); one MLMG "smooth" = applyBC + red
             half-sweep, applyBC + black half-sweep;
  transfer   restriction = mean of the 8 fine cells, interpolation = piecewise constant (MLCellLinOp::restriction /
             interpolation, ratio 2), coarse b = mean of the 4 coincident fine faces (average_down_faces);
  cycle      MLMG::solve / oneIter / mgVcycle with nu1 = nu2 = 2 (mac_proj.num_pre_smooth / num_post_smooth), bottom
             BiCGStab (MLCGSolver, bottom_rtol 1e-4, maxiter 200; falls back to smoothing), solvability offsets for
             all-Neumann / periodic problems (mean of rhs removed, also at the bottom level);
  projection rhs = -div(u_mac); solve; u_mac += flux, flux = -b grad phi with the BC ghost cells (MacProjector::project).

It is pinned by an independent SciPy assembly + sparse direct solve of the same discrete problem
(tests/golden/make_golden_mac.py) and by algebraic identities (tests/test_mac_oracle.py).
Arrays: cells (nz, ny, nx); x faces (nz, ny, nx+1), y faces (nz, ny+1, nx), z faces (nz+1, ny, nx); C order = i fastest.
"""
import numpy as np

PER, NEU, DIR = 0, 1, 2
OMEGA = 1.15
MAXORDER = 3


class Params:
    def __init__(self, n, dx, bclo, bchi, max_coarsening_level=100, maxiter=200, nu1=2, nu2=2, bottom_maxiter=200, bottom_rtol=1e-4,
                 bottom_atol=-1.0, verbose=0):
        self.n, self.dx = tuple(int(x) for x in n), tuple(float(x) for x in dx)
        self.bclo, self.bchi = tuple(int(x) for x in bclo), tuple(int(x) for x in bchi)
        self.max_coarsening_level, self.maxiter, self.nu1, self.nu2 = max_coarsening_level, maxiter, nu1, nu2
        self.bottom_maxiter, self.bottom_rtol, self.bottom_atol, self.verbose = bottom_maxiter, bottom_rtol, bottom_atol, verbose
        for d in range(3):
            assert (self.bclo[d] == PER) == (self.bchi[d] == PER)


AX = {0: 2, 1: 1, 2: 0}   # direction -> numpy axis of a cell array


def _sl(ax, s):
    idx = [slice(None)] * 3
    idx[ax] = s
    return tuple(idx)


class Level:
    def __init__(self, n, dx, b):
        self.n, self.dx, self.b = n, dx, b                    # b: [bx, by, bz]
        self.dh = [1.0 / (dx[d] * dx[d]) for d in range(3)]


class MG:
    def __init__(self, p, beta):
        """beta: [bx, by, bz] face arrays, or a float (MacProjector::initProjector(ba, dm, info, const_beta))"""
        self.p = p
        n = p.n
        if np.isscalar(beta):
            beta = [np.full((n[2], n[1], n[0] + 1), float(beta)), np.full((n[2], n[1] + 1, n[0]), float(beta)),
                    np.full((n[2] + 1, n[1], n[0]), float(beta))]
        self.lv = [Level(n, p.dx, [np.array(x, dtype=np.float64) for x in beta])]
        lev = 0
        while lev < p.max_coarsening_level and all(m % 2 == 0 and m // 2 >= 2 for m in self.lv[-1].n):
            f = self.lv[-1]
            nc = tuple(m // 2 for m in f.n)
            bx, by, bz = f.b
            cbx = 0.25 * (bx[0::2, 0::2, 0::2] + bx[1::2, 0::2, 0::2] + bx[0::2, 1::2, 0::2] + bx[1::2, 1::2, 0::2])
            cby = 0.25 * (by[0::2, 0::2, 0::2] + by[1::2, 0::2, 0::2] + by[0::2, 0::2, 1::2] + by[1::2, 0::2, 1::2])
            cbz = 0.25 * (bz[0::2, 0::2, 0::2] + bz[0::2, 1::2, 0::2] + bz[0::2, 0::2, 1::2] + bz[0::2, 1::2, 1::2])
            self.lv.append(Level(nc, tuple(2 * h for h in f.dx), [cbx, cby, cbz]))
            lev += 1
        self.singular = all(b != DIR for b in p.bclo + p.bchi)
        self.bottom_iters = 0

    # -- boundary conditions -------------------------------------------------------------------------------------
    def _ghosts(self, L, phi, d):
        """(lo ghost, hi ghost) planes of phi in direction d (homogeneous BC values)"""
        ax = AX[d]
        n = L.n[d]

        def one(bc, first, second, wrap):
            if bc == PER:
                return phi[_sl(ax, wrap)]
            if bc == NEU:
                return phi[_sl(ax, first)]
            nxo = min(n + 1, MAXORDER)
            if nxo >= 3:
                return -2.0 * phi[_sl(ax, first)] + phi[_sl(ax, second)] / 3.0
            return -phi[_sl(ax, first)]
        return one(self.p.bclo[d], 0, 1, n - 1), one(self.p.bchi[d], n - 1, n - 2, 0)

    def _bc_coef(self, L, d):
        """coefficient of the first interior cell in the ghost-cell formula (m_undrrelxr / mllinop_comp_interp_coef0)"""
        def one(bc):
            if bc == PER:
                return 0.0
            if bc == NEU:
                return 1.0
            return -2.0 if min(L.n[d] + 1, MAXORDER) >= 3 else -1.0
        return one(self.p.bclo[d]), one(self.p.bchi[d])

    def _neighbours(self, L, phi, d):
        """phi(i-1), phi(i+1) along d with BC ghost cells"""
        ax = AX[d]
        glo, ghi = self._ghosts(L, phi, d)
        lo = np.concatenate([np.expand_dims(glo, ax), phi[_sl(ax, slice(0, -1))]], axis=ax)
        hi = np.concatenate([phi[_sl(ax, slice(1, None))], np.expand_dims(ghi, ax)], axis=ax)
        return lo, hi

    def _faces(self, L, d):
        ax = AX[d]
        b = L.b[d]
        return b[_sl(ax, slice(0, -1))], b[_sl(ax, slice(1, None))]

    # -- operator ------------------------------------------------------------------------------------------------
    def adotx(self, lev, phi):
        L = self.lv[lev]
        y = np.zeros_like(phi)
        for d in range(3):
            lo, hi = self._neighbours(L, phi, d)
            bl, bh = self._faces(L, d)
            y -= L.dh[d] * (bh * (hi - phi) - bl * (phi - lo))
        return y

    def residual(self, lev, phi, rhs):
        return rhs - self.adotx(lev, phi)

    def smooth(self, lev, phi, rhs, ncalls=1):
        """ncalls MLMG smooth calls: each = red half-sweep then black half-sweep (abec_gsrb), in place"""
        L = self.lv[lev]
        k, j, i = np.meshgrid(np.arange(L.n[2]), np.arange(L.n[1]), np.arange(L.n[0]), indexing="ij")
        par = (i + j + k) % 2
        gamma = np.zeros_like(phi)
        delta = np.zeros_like(phi)
        for d in range(3):
            ax = AX[d]
            bl, bh = self._faces(L, d)
            gamma += L.dh[d] * (bl + bh)
            clo, chi = self._bc_coef(L, d)
            first, last = _sl(ax, 0), _sl(ax, L.n[d] - 1)
            delta[first] += L.dh[d] * bl[first] * clo
            delta[last] += L.dh[d] * bh[last] * chi
        for _ in range(ncalls):
            for redblack in (0, 1):
                rho = np.zeros_like(phi)
                for d in range(3):
                    lo, hi = self._neighbours(L, phi, d)
                    bl, bh = self._faces(L, d)
                    rho += L.dh[d] * (bl * lo + bh * hi)
                res = rhs - (gamma * phi - rho)
                sel = ((par + redblack) % 2) == 0
                phi[sel] += (OMEGA / (gamma - delta) * res)[sel]
        return phi

    @staticmethod
    def restrict(fine):
        return 0.125 * (fine[0::2, 0::2, 0::2] + fine[1::2, 0::2, 0::2] + fine[0::2, 1::2, 0::2] + fine[1::2, 1::2, 0::2] +
                        fine[0::2, 0::2, 1::2] + fine[1::2, 0::2, 1::2] + fine[0::2, 1::2, 1::2] + fine[1::2, 1::2, 1::2])

    @staticmethod
    def interp_add(fine, crse):
        fine += np.repeat(np.repeat(np.repeat(crse, 2, axis=0), 2, axis=1), 2, axis=2)
        return fine

    # -- bottom solver: MLCGSolver::solve_bicgstab, plain dot products, homogeneous BCs ------------------------------
    def bottom_solve(self, x, b):
        lev = len(self.lv) - 1
        p = self.p
        if self.singular:
            b = b - b.mean()
        ret, it = self._bicgstab(lev, x, b)
        self.bottom_iters += it
        if ret != 0:                      # MLMG::bottomSolve: on failure start over with smoothing
            x[...] = 0.0
            self.smooth(lev, x, b, 8)     # nuf = 8 smooth calls
        return x

    def _bicgstab(self, lev, sol, rhs):
        p = self.p
        eps_rel, eps_abs = p.bottom_rtol, p.bottom_atol
        dot = lambda a, b: float((a * b).sum())
        ninf = lambda a: float(np.abs(a).max()) if a.size else 0.0
        sorig = sol.copy()
        r = rhs - self.adotx(lev, sol)
        rh = r.copy()
        sol[...] = 0.0
        rnorm = ninf(r)
        rnorm0 = rnorm
        if rnorm0 == 0 or rnorm0 < eps_abs:
            sol += sorig
            return 0, 0
        rho_1 = alpha = omega = 0.0
        pvec = np.zeros_like(r); v = np.zeros_like(r)
        ret, it = 0, 0
        for it in range(1, p.bottom_maxiter + 1):
            rho = dot(rh, r)
            if rho == 0:
                ret = 1; break
            if it == 1:
                pvec = r.copy()
            else:
                beta = (rho / rho_1) * (alpha / omega)
                pvec = r + beta * (pvec - omega * v)
            v = self.adotx(lev, pvec)
            rhTv = dot(rh, v)
            if rhTv == 0:
                ret = 2; break
            alpha = rho / rhTv
            sol += alpha * pvec
            s = r - alpha * v
            rnorm = ninf(s)
            if rnorm < eps_rel * rnorm0 or rnorm < eps_abs:
                r = s
                break
            t = self.adotx(lev, s)
            tt = dot(t, t)
            if tt == 0:
                ret = 3; break
            omega = dot(t, s) / tt
            sol += omega * s
            r = s - omega * t
            rnorm = ninf(r)
            if rnorm < eps_rel * rnorm0 or rnorm < eps_abs:
                break
            if omega == 0:
                ret = 4; break
            rho_1 = rho
        else:
            ret = 8
        if ret == 0 and not (rnorm < eps_rel * rnorm0 or rnorm < eps_abs):
            ret = 8
        if ret == 0 or ret == 8:
            sol += sorig
        else:
            sol[...] = sorig
        return (0 if ret == 0 else ret), it

    # -- MLMG ------------------------------------------------------------------------------------------------------
    def vcycle(self, res0):
        p = self.p
        nl = len(self.lv)
        res = [res0] + [None] * (nl - 1)
        cor = [None] * nl
        for l in range(nl - 1):
            cor[l] = np.zeros_like(res[l])
            self.smooth(l, cor[l], res[l], p.nu1)
            res[l + 1] = self.restrict(res[l] - self.adotx(l, cor[l]))
        cor[nl - 1] = np.zeros_like(res[nl - 1])
        if nl == 1:
            self.bottom_solve(cor[0], res[0])
        else:
            self.bottom_solve(cor[nl - 1], res[nl - 1])
        for l in range(nl - 2, -1, -1):
            self.interp_add(cor[l], cor[l + 1])
            self.smooth(l, cor[l], res[l], p.nu2)
        return cor[0]

    def solve(self, phi, rhs, rtol, atol):
        """MLMG::solve on level-0 arrays; phi is the initial guess (in/out).  Returns a stats dict."""
        p = self.p
        rhs = rhs.copy()
        if self.singular:
            rhs -= rhs.mean()
        rhsnorm = float(np.abs(rhs).max())
        res = self.residual(0, phi, rhs)
        resnorm0 = float(np.abs(res).max())
        maxnorm = max(rhsnorm, resnorm0)
        target = max(atol, max(rtol, 1e-16) * maxnorm)
        st = dict(iters=0, status=0, rhsnorm=rhsnorm, resnorm0=resnorm0, resnorm=resnorm0, nlevels=len(self.lv))
        if resnorm0 <= target:
            return st
        for it in range(p.maxiter):
            phi += self.vcycle(res)
            res = self.residual(0, phi, rhs)
            st["resnorm"] = float(np.abs(res).max())
            st["iters"] = it + 1
            if p.verbose:
                print(f"MLMG: Iteration {it + 1:3d} Fine resid/bnorm = {st['resnorm'] / maxnorm:.12g}")
            if st["resnorm"] <= target:
                break
            if not st["resnorm"] <= 1e20 * maxnorm:
                st["status"] = 2
                break
        else:
            st["status"] = 1
        st["bottom_iters"] = self.bottom_iters
        return st

    # -- MacProjector ------------------------------------------------------------------------------------------------
    def divergence(self, u, v, w):
        dx = self.p.dx
        return (u[:, :, 1:] - u[:, :, :-1]) / dx[0] + (v[:, 1:, :] - v[:, :-1, :]) / dx[1] + (w[1:] - w[:-1]) / dx[2]

    def fluxes(self, phi):
        """-b grad phi on all faces of level 0, BC ghost cells included (MLMG::getFluxes, Location::FaceCenter)"""
        L = self.lv[0]
        out = []
        for d in range(3):
            ax = AX[d]
            glo, ghi = self._ghosts(L, phi, d)
            ext = np.concatenate([np.expand_dims(glo, ax), phi, np.expand_dims(ghi, ax)], axis=ax)
            g = (ext[_sl(ax, slice(1, None))] - ext[_sl(ax, slice(0, -1))]) / L.dx[d]
            out.append(-L.b[d] * g)
        return out


def project(p, umac, vmac, wmac, beta, rtol, atol, phi0=None):
    """MacProjector::project(rtol, atol) (phi0 is None: start from 0) / project(mac_phi, rtol, atol).  Face velocities are
    modified in place.  Returns dict(phi, stats)."""
    mg = MG(p, beta)
    rhs = -mg.divergence(umac, vmac, wmac)
    phi = np.zeros(rhs.shape) if phi0 is None else np.array(phi0, dtype=np.float64)
    st = mg.solve(phi, rhs, rtol, atol)
    fx, fy, fz = mg.fluxes(phi)
    umac += fx; vmac += fy; wmac += fz
    return dict(phi=phi, stats=st, mg=mg)
