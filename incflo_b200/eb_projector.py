"""Host-side mirror of Hydro::NodalProjector built with an EBFArrayBoxFactory, as incflo drives it under AMREX_USE_EB
(src/projection/incflo_apply_nodal_projection.cpp:130-136, :181-201, :215-266), backed by libb200np.so (b200eb_* of
include/b200np.h).  Arrays: numpy (host, staged inside the call) or torch CUDA tensors (zero copy), C-contiguous float64,
(ncomp, nz, ny, nx) -- amrex::Array4 order.  Nodal arrays: (nz+1, ny+1, nx+1), the box [0, n_cell]."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import EBFlow, Geom, Stats
from .nodal_projector import ProjectionError, _ptr_box, nodal_proj_opts

OP_SMOOTH, OP_RESIDUAL, OP_RESTRICT, OP_INTERP, OP_BOTTOM, OP_APPLY = range(6)


def eb_flow(has_normal=False, normal=(0.0, 0.0, 0.0), normal_tol=0.0, vel_mag=None, velocity=(0.0, 0.0, 0.0), density=1.0, tracer=()):
    """the eb_flow.* inputs (incflo's EBFlow struct read in src/setup/init.cpp)"""
    f = EBFlow()
    f.has_normal = int(bool(has_normal))
    for d in range(3):
        f.normal[d] = float(normal[d]); f.velocity[d] = float(velocity[d])
    f.normal_tol = float(normal_tol)
    f.is_mag = int(vel_mag is not None)
    f.vel_mag = float(vel_mag) if vel_mag is not None else 0.0
    f.density = float(density)
    f.ntrac = len(tracer)
    for m, t in enumerate(tracer):
        f.tracer[m] = float(t)
    return f


class EBNodalProjector:
    """NodalProjector(vel, sigma | const_sigma, geom, LPInfo) with an EB factory: setDomainBC happens in the ctor (bclo, bchi =
    LinOpBCType codes of incflo::get_projection_bc)."""

    def __init__(self, n_cell, dx, bclo, bchi, vfrac, intg, opts=None, device=0):
        self._L = _lib.lib()
        self.n = tuple(int(x) for x in n_cell)
        g = Geom()
        for d in range(3):
            g.n_cell[d] = self.n[d]; g.dx[d] = float(dx[d]); g.bc_lo[d] = int(bclo[d]); g.bc_hi[d] = int(bchi[d])
        self.opts = opts if opts is not None else nodal_proj_opts()
        h = C.c_void_p()
        rc = self._L.b200eb_create(C.byref(h), C.byref(g), C.byref(self.opts), device)
        if rc != 0:
            raise ProjectionError(rc)
        self._h = h
        self.stats = Stats()
        self.set_geometry(vfrac, intg)

    def _chk(self, rc):
        if rc != 0:
            raise ProjectionError(rc)

    def set_geometry(self, vfrac, intg):
        pv, bv, _ = _ptr_box(vfrac, (0, 0, 0), 1)
        pi, bi, _ = _ptr_box(intg, (0, 0, 0), 18)
        self._chk(self._L.b200eb_set_geometry(self._h, pv, C.byref(bv), pi, C.byref(bi)))

    def setEBInflowVelocity(self, eb_vel, bnorm=None, bintg=None):
        """getLinOp().setEBInflowVelocity(lev, eb_vel); None clears it"""
        if eb_vel is None:
            self._chk(self._L.b200eb_set_eb_inflow_velocity(self._h, None, None, None, None, None, None))
            return
        pv, bv, _ = _ptr_box(eb_vel, (0, 0, 0), 3)
        pn, bn, _ = _ptr_box(bnorm, (0, 0, 0), 3)
        pb, bb, _ = _ptr_box(bintg, (0, 0, 0), 8)
        self._chk(self._L.b200eb_set_eb_inflow_velocity(self._h, pv, C.byref(bv), pn, C.byref(bn), pb, C.byref(bb)))

    def set_eb_flow(self, flow, bnorm, nghost, eb_vel=None, eb_density=None, eb_tracer=None):
        """incflo::set_eb_velocity / set_eb_density / set_eb_tracer: outputs have nghost ghost cells, (ncomp, nz+2ng, ny+2ng, nx+2ng)"""
        pn, bn, _ = _ptr_box(bnorm, (0, 0, 0), 3)
        lo = (-nghost,) * 3
        pv, bv, _ = _ptr_box(eb_vel, lo, 3)
        pd, bd, _ = _ptr_box(eb_density, lo, 1)
        pt, bt, _ = _ptr_box(eb_tracer, lo, max(flow.ntrac, 1))
        ref = lambda b: C.byref(b) if b is not None else None
        self._chk(self._L.b200eb_set_eb_flow(self._h, C.byref(flow), int(nghost), pn, C.byref(bn), pv, ref(bv), pd, ref(bd), pt, ref(bt)))

    def project(self, vel, sigma, rtol, atol, phi=None, gphi=None, ng=1):
        """vel (3, nz+2ng, ny+2ng, nx+2ng) in/out; sigma: cell array or float"""
        pv, bv, _ = _ptr_box(vel, (-ng,) * 3, 3)
        if np.isscalar(sigma):
            ps, bs, cs = None, None, float(sigma)
        else:
            ps, bs, _ = _ptr_box(sigma, (0, 0, 0), 1)
            cs = 0.0
        pp, bp, _ = _ptr_box(phi, (0, 0, 0), 1)
        pg, bg, _ = _ptr_box(gphi, (0, 0, 0), 3)
        ref = lambda b: C.byref(b) if b is not None else None
        _lib.torch_sync()
        rc = self._L.b200eb_project(self._h, pv, C.byref(bv), ps, ref(bs), cs, pp, ref(bp), pg, ref(bg), float(rtol), float(atol), C.byref(self.stats))
        self._chk(rc)
        return self.stats

    def apply_nodal_projection(self, velocity, velocity_o, density, ro_0, gp, p_nd, scaling_factor, incremental, proj_for_small_dt, rtol, atol,
                               inflow_vel=None, ng=1):
        """incflo::ApplyNodalProjection(density, time, scaling_factor, incremental) under AMREX_USE_EB"""
        lo = (-ng,) * 3
        pv, bv, _ = _ptr_box(velocity, lo, 3)
        po, _, _ = _ptr_box(velocity_o, lo, 3)
        pr, br, _ = _ptr_box(density, (0, 0, 0), 1)
        pg, bg, _ = _ptr_box(gp, (0, 0, 0), 3)
        pp, bp, _ = _ptr_box(p_nd, (0, 0, 0), 1)
        pi, _, _ = _ptr_box(inflow_vel, lo, 3)
        ref = lambda b: C.byref(b) if b is not None else None
        _lib.torch_sync()
        rc = self._L.b200eb_apply_nodal_projection(self._h, pv, C.byref(bv), po, pr, ref(br), float(ro_0), pg, C.byref(bg), pp, C.byref(bp), pi,
                                                   float(scaling_factor), int(incremental), int(proj_for_small_dt), float(rtol), float(atol),
                                                   C.byref(self.stats))
        self._chk(rc)
        return self.stats

    # -- multi-box MultiFabs (nodal_projector.MultiFab: amr.max_grid_size < domain) --
    def set_geometry_mf(self, vfrac, intg):
        self._chk(self._L.b200eb_set_geometry_mf(self._h, vfrac.ref(), intg.ref()))

    def project_mf(self, vel, sigma, rtol, atol, phi=None, gphi=None):
        """vel / phi / gphi (and sigma unless it is a float): MultiFab"""
        cs = float(sigma) if np.isscalar(sigma) else 0.0
        ref = lambda m: m.ref() if m is not None else None
        _lib.torch_sync()
        rc = self._L.b200eb_project_mf(self._h, vel.ref(), None if np.isscalar(sigma) else sigma.ref(), cs, ref(phi), ref(gphi), float(rtol), float(atol),
                                       C.byref(self.stats))
        self._chk(rc)
        return self.stats

    def apply_nodal_projection_mf(self, velocity, velocity_o, density, ro_0, gp, p_nd, scaling_factor, incremental, proj_for_small_dt, rtol, atol,
                                  inflow_vel=None):
        ref = lambda m: m.ref() if m is not None else None
        _lib.torch_sync()
        rc = self._L.b200eb_apply_nodal_projection_mf(self._h, velocity.ref(), ref(velocity_o), ref(density), float(ro_0), gp.ref(), p_nd.ref(), ref(inflow_vel),
                                                      float(scaling_factor), int(incremental), int(proj_for_small_dt), float(rtol), float(atol),
                                                      C.byref(self.stats))
        self._chk(rc)
        return self.stats

    # -- per-kernel hooks used by the parity tests (host numpy arrays, natural node order) --
    def nlevels(self):
        return self._L.b200eb_nlevels(self._h)

    def level_dims(self, lev):
        n, nn = (C.c_int * 3)(), (C.c_int * 3)()
        self._chk(self._L.b200eb_level_dims(self._h, lev, C.byref(n), C.byref(nn)))
        return tuple(n), tuple(nn)

    def build_stencils(self, sigma):
        if np.isscalar(sigma):
            self._chk(self._L.b200eb_build_stencils(self._h, None, None, float(sigma)))
        else:
            ps, bs, _ = _ptr_box(sigma, (0, 0, 0), 1)
            self._chk(self._L.b200eb_build_stencils(self._h, ps, C.byref(bs), 0.0))

    def level_stencil(self, lev):
        _, nn = self.level_dims(lev)
        out = np.empty((14, nn[2], nn[1], nn[0]))
        self._chk(self._L.b200eb_level_stencil(self._h, lev, C.c_void_p(out.ctypes.data)))
        return out

    def level_op(self, lev, op, arg=0, a=None, b=None, out_lev=None):
        _, nn = self.level_dims(lev if out_lev is None else out_lev)
        out = np.empty((nn[2], nn[1], nn[0]))
        ka = None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        kb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        pa = None if ka is None else C.c_void_p(ka.ctypes.data)
        pb = None if kb is None else C.c_void_p(kb.ctypes.data)
        self._chk(self._L.b200eb_level_op(self._h, lev, op, arg, pa, pb, C.c_void_p(out.ctypes.data)))
        return out

    def compute_rhs(self, vel, ng=1):
        pv, bv, _ = _ptr_box(vel, (-ng,) * 3, 3)
        _, nn = self.level_dims(0)
        out = np.empty((nn[2], nn[1], nn[0]))
        self._chk(self._L.b200eb_compute_rhs(self._h, pv, C.byref(bv), C.c_void_p(out.ctypes.data)))
        return out

    def time_op(self, lev, op, arg=1, reps=10):
        ms = C.c_double()
        self._chk(self._L.b200eb_time_op(self._h, lev, op, arg, reps, C.byref(ms)))
        return ms.value

    def close(self):
        if self._h is not None:
            self._L.b200eb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
