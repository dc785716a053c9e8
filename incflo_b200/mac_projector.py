"""Host-side mirror of Hydro::MacProjector as incflo uses it
(src/convection/incflo_compute_MAC_projected_velocities.cpp:69-129, :280-299), backed by libb200np.so (b200mac_* of
include/b200np.h).  Arrays: numpy (host, staged inside the call) or torch CUDA tensors (zero copy), C-contiguous float64;
x faces (nz, ny, nx+1), y faces (nz, ny+1, nx), z faces (nz+1, ny, nx), cells (nz, ny, nx)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FabBox, Geom, Stats
from .nodal_projector import ProjectionError, nodal_proj_opts

OP_SMOOTH, OP_RESIDUAL, OP_RESTRICT, OP_INTERP, OP_BOTTOM = range(5)


def mac_proj_opts(**keys):
    """mac_proj.* keys (src/setup/init.cpp:165-170 + Hydro::MacProjector::setOptions); MLMG defaults maxiter = bottom_maxiter = 200"""
    base = dict(maxiter=200, bottom_maxiter=200)
    base.update(keys)
    return nodal_proj_opts(**base)


def _ptr_box(a):
    if a is None:
        return None, None
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        ptr, shape = C.c_void_p(a.data_ptr()), tuple(a.shape)
    else:
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        ptr, shape = C.c_void_p(a.ctypes.data), a.shape
    b = FabBox()
    for d in range(3):
        b.lo[d] = 0; b.hi[d] = shape[2 - d] - 1
    b.ncomp = 1
    return ptr, b


class MacProjector:
    """macproj->initProjector(lp_info, inv_rho | const beta); setDomainBC(lo, hi); project(rtol, atol)"""

    def __init__(self, n_cell, dx, bclo, bchi, opts=None, device=0):
        self._L = _lib.lib()
        self.n = tuple(int(x) for x in n_cell)
        g = Geom()
        for d in range(3):
            g.n_cell[d] = self.n[d]; g.dx[d] = float(dx[d]); g.bc_lo[d] = int(bclo[d]); g.bc_hi[d] = int(bchi[d])
        self.opts = opts if opts is not None else mac_proj_opts()
        h = C.c_void_p()
        rc = self._L.b200mac_create(C.byref(h), C.byref(g), C.byref(self.opts), device)
        if rc != 0:
            raise ProjectionError(rc)
        self._h = h
        self.stats = Stats()

    def updateCoeffs(self, beta):
        """beta: (bx, by, bz) face arrays dt / rho (initProjector / updateCoeffs), or a float (updateBeta)"""
        if np.isscalar(beta):
            rc = self._L.b200mac_set_coeffs(self._h, None, None, None, None, None, None, float(beta))
        else:
            (px, bx), (py, by), (pz, bz) = (_ptr_box(a) for a in beta)
            rc = self._L.b200mac_set_coeffs(self._h, px, C.byref(bx), py, C.byref(by), pz, C.byref(bz), 0.0)
        if rc != 0:
            raise ProjectionError(rc)

    def project(self, umac, vmac, wmac, rtol, atol, mac_phi=None, use_phi_as_guess=False):
        (pu, bu), (pv, bv), (pw, bw) = _ptr_box(umac), _ptr_box(vmac), _ptr_box(wmac)
        pp, bp = _ptr_box(mac_phi)
        rc = self._L.b200mac_project(self._h, pu, C.byref(bu), pv, C.byref(bv), pw, C.byref(bw), pp, C.byref(bp) if bp is not None else None,
                                     int(use_phi_as_guess), float(rtol), float(atol), C.byref(self.stats))
        if rc != 0:
            raise ProjectionError(rc)
        return self.stats

    # -- per-kernel hooks used by the parity tests (host numpy arrays) --
    def nlevels(self):
        return self._L.b200mac_nlevels(self._h)

    def level_dims(self, lev):
        n = (C.c_int * 3)()
        self._L.b200mac_level_dims(self._h, lev, C.byref(n))
        return tuple(n)

    def level_op(self, lev, op, arg=0, a=None, b=None, out_lev=None):
        n = self.level_dims(lev if out_lev is None else out_lev)
        out = np.empty((n[2], n[1], n[0]))
        p = lambda x: None if x is None else C.c_void_p(np.ascontiguousarray(x, dtype=np.float64).ctypes.data)
        keep = [np.ascontiguousarray(x, dtype=np.float64) for x in (a, b) if x is not None]   # keep the converted arrays alive
        pa = None if a is None else C.c_void_p(keep[0].ctypes.data)
        pb = None if b is None else C.c_void_p(keep[-1].ctypes.data)
        rc = self._L.b200mac_level_op(self._h, lev, op, arg, pa, pb, C.c_void_p(out.ctypes.data))
        if rc != 0:
            raise ProjectionError(rc)
        return out

    def close(self):
        if self._h is not None:
            self._L.b200mac_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
