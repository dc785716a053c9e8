"""Host-side mirror of Hydro::MacProjector as incflo uses it
(src/convection/incflo_compute_MAC_projected_velocities.cpp:69-129, :280-299), backed by libb200np.so (b200mac_* of
include/b200np.h).  Arrays: numpy (host, staged inside the call) or torch CUDA tensors (zero copy), C-contiguous float64;
x faces (nz, ny, nx+1), y faces (nz, ny+1, nx), z faces (nz+1, ny, nx), cells (nz, ny, nx)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FabBox, Geom, MFab, Stats
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


class FaceMultiFab:
    """A face-centred (d = 0, 1, 2) or cell-centred (d = -1) one-component amrex::MultiFab as one rank sees it: boxes = valid CELL
    boxes (lo, hi), arrays = one per box over the cell box grown by ngrow, plus the far face in direction d -- shaped (nz, ny, nx) of
    that allocated box; numpy (host) or torch CUDA tensors."""

    def __init__(self, boxes, arrays, ngrow, d):
        self.boxes, self.arrays, self.ngrow, self.d = list(boxes), list(arrays), int(ngrow), int(d)
        n = len(boxes)
        self._box = (FabBox * n)()
        self._ptr = (C.c_void_p * n)()
        for f, ((lo, hi), a) in enumerate(zip(boxes, arrays)):
            for q in range(3):
                self._box[f].lo[q] = int(lo[q]) - self.ngrow
                self._box[f].hi[q] = int(hi[q]) + self.ngrow + (1 if q == self.d else 0)
                assert tuple(a.shape)[2 - q] == self._box[f].hi[q] - self._box[f].lo[q] + 1
            self._box[f].ncomp = 1
            self._ptr[f] = a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        self.c = MFab(n, self.ngrow, 1, C.cast(self._box, C.POINTER(FabBox)), C.cast(self._ptr, C.POINTER(C.c_void_p)))

    def ref(self):
        return C.byref(self.c)

    @staticmethod
    def split(full, n_cell, max_grid, ngrow, d, to=None, fill=0.0):
        """chop a dense array over the whole domain (faces of direction d: extent n + 1 there) into boxes of at most max_grid cells per
        direction; ghost faces / cells hold `fill` (the projector must neither read nor write them)"""
        nx, ny, nz = n_cell
        boxes, arrs = [], []
        for k0 in range(0, nz, max_grid):
            for j0 in range(0, ny, max_grid):
                for i0 in range(0, nx, max_grid):
                    lo = (i0, j0, k0)
                    hi = (min(i0 + max_grid, nx) - 1, min(j0 + max_grid, ny) - 1, min(k0 + max_grid, nz) - 1)
                    ext = [hi[q] - lo[q] + 1 + (1 if q == d else 0) for q in range(3)]
                    b = np.full((ext[2] + 2 * ngrow, ext[1] + 2 * ngrow, ext[0] + 2 * ngrow), fill)
                    b[ngrow:ngrow + ext[2], ngrow:ngrow + ext[1], ngrow:ngrow + ext[0]] = full[k0:k0 + ext[2], j0:j0 + ext[1], i0:i0 + ext[0]]
                    boxes.append((lo, hi))
                    arrs.append(np.ascontiguousarray(b) if to is None else to(np.ascontiguousarray(b)))
        return FaceMultiFab(boxes, arrs, ngrow, d)

    def assemble(self, n_cell):
        """the valid regions put back together (shared faces: the later box wins -- they agree)"""
        ext = [n_cell[q] + (1 if q == self.d else 0) for q in range(3)]
        out = np.zeros((ext[2], ext[1], ext[0]))
        g = self.ngrow
        for (lo, hi), a in zip(self.boxes, self.arrays):
            a = a.detach().cpu().numpy() if hasattr(a, "detach") else a
            e = [hi[q] - lo[q] + 1 + (1 if q == self.d else 0) for q in range(3)]
            out[lo[2]:lo[2] + e[2], lo[1]:lo[1] + e[1], lo[0]:lo[0] + e[0]] = a[g:g + e[2], g:g + e[1], g:g + e[0]]
        return out


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
            _lib.torch_sync()
            rc = self._L.b200mac_set_coeffs(self._h, None, None, None, None, None, None, float(beta))
        else:
            (px, bx), (py, by), (pz, bz) = (_ptr_box(a) for a in beta)
            _lib.torch_sync()
            rc = self._L.b200mac_set_coeffs(self._h, px, C.byref(bx), py, C.byref(by), pz, C.byref(bz), 0.0)
        if rc != 0:
            raise ProjectionError(rc)

    def project(self, umac, vmac, wmac, rtol, atol, mac_phi=None, use_phi_as_guess=False):
        (pu, bu), (pv, bv), (pw, bw) = _ptr_box(umac), _ptr_box(vmac), _ptr_box(wmac)
        pp, bp = _ptr_box(mac_phi)
        _lib.torch_sync()
        rc = self._L.b200mac_project(self._h, pu, C.byref(bu), pv, C.byref(bv), pw, C.byref(bw), pp, C.byref(bp) if bp is not None else None,
                                     int(use_phi_as_guess), float(rtol), float(atol), C.byref(self.stats))
        if rc != 0:
            raise ProjectionError(rc)
        return self.stats

    def updateCoeffs_mf(self, bx, by, bz):
        """initProjector / updateCoeffs over multi-box face MultiFabs (FaceMultiFab)"""
        _lib.torch_sync()
        rc = self._L.b200mac_set_coeffs_mf(self._h, bx.ref(), by.ref(), bz.ref())
        if rc != 0:
            raise ProjectionError(rc)

    def project_mf(self, umac, vmac, wmac, rtol, atol, mac_phi=None, use_phi_as_guess=False):
        _lib.torch_sync()
        rc = self._L.b200mac_project_mf(self._h, umac.ref(), vmac.ref(), wmac.ref(), mac_phi.ref() if mac_phi is not None else None,
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
