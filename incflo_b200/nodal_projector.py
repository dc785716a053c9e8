"""Host-side mirror of the reference interface for the nodal-projection path.

Same names / argument meaning / error behaviour as the reference:
  * ``NodalProjector``  <->  Hydro::NodalProjector as used by incflo
    (src/projection/incflo_apply_nodal_projection.cpp:181-219): ctor(vel, sigma |
    const_sigma, geom, LPInfo), setDomainBC(lo, hi), project(rtol, atol), getPhi(),
    getGradPhi().
  * ``apply_nodal_projection``  <->  incflo::ApplyNodalProjection (:29-267).
  * ``get_projection_bc``  <->  incflo::get_projection_bc
    (src/projection/incflo_projection_bc.cpp:5-41).
Everything numerical happens in libb200np.so (sm_100a CUDA) through the C ABI of
include/b200np.h.  Arrays are numpy (host, staged inside the call) or torch CUDA
tensors (zero copy), shape (ncomp, nz, ny, nx) / (nz, ny, nx), C-contiguous float64,
which is amrex::Array4 order (i fastest, component outermost).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FabBox, Geom, MFab, Opts, Stats

BC_PERIODIC, BC_NEUMANN, BC_DIRICHLET, BC_INFLOW = 0, 1, 2, 3
A_SOL, A_RHS, A_RES, A_COR, A_RESCOR, A_SIGMA = range(6)
OP_SMOOTH, OP_RESIDUAL, OP_RESTRICT, OP_INTERP, OP_BOTTOM, OP_VCYCLE, OP_COARSEN_SIGMA = range(7)
FACE_DEFAULT, FACE_DIRECTION_DEPENDENT, FACE_MIXED = 0, 1, 2   # enum b200np_face_type
SMOOTH_ZERO_START = 0x10000   # B200NP_SMOOTH_ZERO_START: OP_SMOOTH as the V-cycle's "cor = 0" pre-smooth

# incflo BC names (src/boundary_conditions/boundary_conditions.cpp:30-222) -> LinOpBCType
_INCFLO_BC = {"pi": BC_DIRICHLET, "pressure_inflow": BC_DIRICHLET, "po": BC_DIRICHLET, "pressure_outflow": BC_DIRICHLET,
              "mi": BC_INFLOW, "mass_inflow": BC_INFLOW, "dd": BC_INFLOW, "direction_dependent": BC_INFLOW,
              "mixed": BC_INFLOW, "sw": BC_NEUMANN, "slip_wall": BC_NEUMANN, "nsw": BC_NEUMANN,
              "no_slip_wall": BC_NEUMANN}


class ProjectionError(RuntimeError):
    """The reference calls amrex::Abort in these situations."""

    def __init__(self, status):
        self.status = status
        super().__init__(_lib.lib().b200np_strerror(status).decode())


def get_projection_bc(is_periodic, bc_types):
    """incflo::get_projection_bc for one side: list of 3 incflo BC names -> LinOpBCType codes."""
    r = []
    for d in range(3):
        if is_periodic[d]:
            r.append(BC_PERIODIC)
        else:
            if bc_types[d] not in _INCFLO_BC:
                raise ProjectionError(3)  # "get_projection_bc: undefined BC type"
            r.append(_INCFLO_BC[bc_types[d]])
    return tuple(r)


def nodal_proj_opts(**keys):
    """nodal_proj.* ParmParse keys -> b200np_opts (defaults: src/incflo.H:449-458, InputsMultigrid.rst)."""
    o = Opts()
    _lib.lib().b200np_default_opts(C.byref(o))
    alias = {"mg_max_coarsening_level": "mg_max_coarsening_level"}
    for k, v in keys.items():
        k = alias.get(k, k)
        if k == "tile":
            for d in range(3):
                o.tile[d] = int(v[d])
        elif k == "bottom_solver" and isinstance(v, str):
            o.bottom_solver = {"bicgcg": 0, "bicgstab": 0, "cg": 0, "cgbicg": 0, "smoother": 1}[v]
        elif k in ("mg_rtol", "mg_atol"):
            continue  # passed to project()
        else:
            if not hasattr(o, k):
                raise KeyError("unknown nodal_proj key: " + k)
            setattr(o, k, v)
    return o


def _ptr_box(a, lo, ncomp):
    """(pointer, FabBox) of a numpy array or torch tensor laid out (ncomp?, nz, ny, nx)."""
    if a is None:
        return None, None, None
    shape = tuple(a.shape)
    if len(shape) == 4:
        assert shape[0] == ncomp, (shape, ncomp)
        nz, ny, nx = shape[1:]
    else:
        assert ncomp == 1
        nz, ny, nx = shape
    b = FabBox()
    for d, (l, n) in enumerate(zip(lo, (nx, ny, nz))):
        b.lo[d] = int(l); b.hi[d] = int(l) + n - 1
    b.ncomp = ncomp
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data, b, a
    # torch tensor
    import torch
    assert a.dtype == torch.float64 and a.is_contiguous()
    return a.data_ptr(), b, a


class NodalProjector:
    """Hydro::NodalProjector, single level, backed by libb200np.so.

    vel: (3, nz+2ng, ny+2ng, nx+2ng); sigma: (nz, ny, nx) or None with const_sigma.
    geom: dict(n_cell=(nx,ny,nz), dx=(..), is_periodic=(..)) -- amrex::Geometry.
    """

    def __init__(self, vel, sigma=None, const_sigma=None, geom=None, ng=None, opts=None, device=0):
        self._L = _lib.lib()
        self.vel, self.sigma, self.const_sigma = vel, sigma, (1.0 if const_sigma is None else float(const_sigma))
        self.n = tuple(int(x) for x in geom["n_cell"])
        self.dx = tuple(float(x) for x in geom["dx"])
        self.per = tuple(bool(x) for x in geom.get("is_periodic", (1, 1, 1)))
        self.ng = int(ng) if ng is not None else (vel.shape[-1] - self.n[0]) // 2
        self.opts = opts if opts is not None else nodal_proj_opts()
        self.device = device
        self.bclo = tuple(BC_PERIODIC if p else BC_NEUMANN for p in self.per)
        self.bchi = self.bclo
        self._h = None
        self._phi = None
        self._gphi = None
        self._faces = None
        self.stats = Stats()

    # -- reference API ----------------------------------------------------------------
    def setOversetMask(self, face_type, mixed_split_dir, mixed_half_num_cells, mask=None):
        """getLinOp().setOversetMask(lev, incflo::make_nodalBC_mask(lev)) for the masks incflo can build: the mixed faces
        (FACE_MIXED per Orientation), the split direction and domain.length(dir) / 2.  `mask` (optional, the caller's
        nodal int32 array) is checked against it; a different mask raises B200NP_ERR_UNSUPPORTED."""
        self._faces = (tuple(int(x) for x in face_type), int(mixed_split_dir), int(mixed_half_num_cells))
        self._mask = mask
        self._destroy()

    def setDomainBC(self, lo, hi):
        self.bclo, self.bchi = tuple(int(x) for x in lo), tuple(int(x) for x in hi)
        self._destroy()

    def project(self, rtol, atol):
        h = self._handle()
        like = self.vel
        self._phi = _empty_like(like, (self.n[2] + 1, self.n[1] + 1, self.n[0] + 1))
        self._gphi = _empty_like(like, (3, self.n[2], self.n[1], self.n[0]))
        pv, bv, _ = _ptr_box(self.vel, (-self.ng,) * 3, 3)
        ps, bs, _ = _ptr_box(self.sigma, (0, 0, 0), 1)
        pp, bp, _ = _ptr_box(self._phi, (0, 0, 0), 1)
        pg, bg, _ = _ptr_box(self._gphi, (0, 0, 0), 3)
        _lib.torch_sync()
        rc = self._L.b200np_project(h, pv, C.byref(bv), ps, C.byref(bs) if bs is not None else None, self.const_sigma,
                                    pp, C.byref(bp), pg, C.byref(bg), float(rtol), float(atol), C.byref(self.stats))
        if rc != 0:
            raise ProjectionError(rc)
        return self.stats

    def getPhi(self):
        return self._phi

    def getGradPhi(self):
        return self._gphi

    # -- plumbing ---------------------------------------------------------------------
    def _handle(self):
        if self._h is None:
            g = Geom()
            for d in range(3):
                g.n_cell[d] = self.n[d]; g.dx[d] = self.dx[d]; g.bc_lo[d] = self.bclo[d]; g.bc_hi[d] = self.bchi[d]
            h = C.c_void_p()
            rc = self._L.b200np_create(C.byref(h), C.byref(g), C.byref(self.opts), self.device)
            if rc != 0:
                raise ProjectionError(rc)
            self._h = h
            if self._faces is not None:
                arr = (C.c_int * 6)(*self._faces[0])
                rc = self._L.b200np_set_face_types(h, C.byref(arr), self._faces[1], self._faces[2])
                if rc != 0:
                    raise ProjectionError(rc)
                if self._mask is not None:
                    m = self._mask
                    ptr = C.c_void_p(m.data_ptr()) if hasattr(m, "data_ptr") else C.c_void_p(m.ctypes.data)
                    b = FabBox()
                    for d in range(3):
                        b.lo[d] = 0; b.hi[d] = tuple(m.shape)[2 - d] - 1
                    b.ncomp = 1
                    rc = self._L.b200np_check_overset_mask(h, ptr, C.byref(b))
                    if rc != 0:
                        raise ProjectionError(rc)
        return self._h

    def _destroy(self):
        if self._h is not None:
            self._L.b200np_destroy(self._h)
            self._h = None

    def close(self):
        self._destroy()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    # -- per-kernel hooks used by the parity tests ---------------------------------------
    def nlevels(self):
        return self._L.b200np_nlevels(self._handle())

    def level_dims(self, lev):
        n = (C.c_int * 3)(); nn = (C.c_int * 3)()
        self._L.b200np_level_dims(self._handle(), lev, C.byref(n), C.byref(nn))
        return tuple(n), tuple(nn)

    def set_sigma(self, sigma=None, const_sigma=1.0):
        ps, bs, _ = _ptr_box(sigma, (0, 0, 0), 1)
        _lib.torch_sync()
        rc = self._L.b200np_set_sigma(self._handle(), ps, C.byref(bs) if bs is not None else None, float(const_sigma))
        if rc:
            raise ProjectionError(rc)

    def level_set(self, lev, which, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        rc = self._L.b200np_level_set(self._handle(), lev, which, arr.ctypes.data)
        if rc:
            raise ProjectionError(rc)

    def level_get(self, lev, which):
        n, nn = self.level_dims(lev)
        shape = (n[2], n[1], n[0]) if which == A_SIGMA else (nn[2], nn[1], nn[0])
        out = np.empty(shape)
        rc = self._L.b200np_level_get(self._handle(), lev, which, out.ctypes.data)
        if rc:
            raise ProjectionError(rc)
        return out

    def level_op(self, lev, op, arg=0):
        rc = self._L.b200np_level_op(self._handle(), lev, op, arg)
        if rc:
            raise ProjectionError(rc)

    def time_op(self, lev, op, arg=1, reps=10):
        ms = C.c_double()
        rc = self._L.b200np_time_op(self._handle(), lev, op, arg, reps, C.byref(ms))
        if rc:
            raise ProjectionError(rc)
        return ms.value


class MultiFab:
    """amrex::MultiFab as one rank sees it: a list of (valid_lo, valid_hi) boxes (cell index space; `nodal` adds the
    high-end nodes) and one array per box, shaped (ncomp, nz, ny, nx) over the box grown by `ngrow` -- numpy (host) or
    torch CUDA tensors.  Mirrors what the C++ side reads off MFIter / fab.box() / fab.dataPtr()."""

    def __init__(self, boxes, arrays, ngrow, ncomp, nodal=False):
        assert len(boxes) == len(arrays) and len(boxes) > 0
        self.boxes, self.arrays, self.ngrow, self.ncomp, self.nodal = list(boxes), list(arrays), int(ngrow), int(ncomp), bool(nodal)
        n = len(boxes)
        self._box = (FabBox * n)()
        self._ptr = (C.c_void_p * n)()
        for f, ((lo, hi), a) in enumerate(zip(boxes, arrays)):
            shape = tuple(a.shape)[-3:]
            for d in range(3):
                self._box[f].lo[d] = int(lo[d]) - self.ngrow
                self._box[f].hi[d] = int(hi[d]) + self.ngrow + (1 if nodal else 0)
                assert shape[2 - d] == self._box[f].hi[d] - self._box[f].lo[d] + 1, (shape, lo, hi)
            self._box[f].ncomp = self.ncomp
            if hasattr(a, "data_ptr"):
                assert a.is_contiguous()
                self._ptr[f] = a.data_ptr()
            else:
                assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
                self._ptr[f] = a.ctypes.data
        self.c = MFab(n, self.ngrow, self.ncomp, C.cast(self._box, C.POINTER(FabBox)), C.cast(self._ptr, C.POINTER(C.c_void_p)))

    def ref(self):
        return C.byref(self.c)

    @staticmethod
    def split(full, n_cell, max_grid, ngrow, ncomp, nodal=False, zlo=0, to=None, origin=(0, 0, 0)):
        """chop a single-box array (ncomp, nz+2ng(+1), ny+.., nx+..) over cells [0,n) x [0,n) x [zlo, zlo+nz) into boxes of
        at most max_grid cells per direction, each with its own ghost frame copied from `full` (zeros where `full` has
        nothing); `to`: callable applied to each box array (e.g. lambda a: torch.from_numpy(a).cuda()); `origin`: index of
        the array's first valid cell (a fine AMR level's box starts at 2 * fine_lo)"""
        nx, ny, nz = n_cell
        ox, oy, oz = origin
        a4 = full.reshape((ncomp,) + tuple(full.shape)[-3:])
        boxes, arrs = [], []
        e = 1 if nodal else 0
        for k0 in range(0, nz, max_grid):
            for j0 in range(0, ny, max_grid):
                for i0 in range(0, nx, max_grid):
                    lo = (i0 + ox, j0 + oy, k0 + zlo + oz)
                    hi = (min(i0 + max_grid, nx) - 1 + ox, min(j0 + max_grid, ny) - 1 + oy, min(k0 + max_grid, nz) - 1 + zlo + oz)
                    sh = (ncomp, hi[2] - lo[2] + 1 + 2 * ngrow + e, hi[1] - lo[1] + 1 + 2 * ngrow + e, hi[0] - lo[0] + 1 + 2 * ngrow + e)
                    b = np.zeros(sh)
                    # the part of the grown box that `full` (cells [-ng, n+ng) around the same origin) holds
                    src = a4[:, k0:k0 + sh[1], j0:j0 + sh[2], i0:i0 + sh[3]]
                    b[:, :src.shape[1], :src.shape[2], :src.shape[3]] = src
                    boxes.append((lo, hi)); arrs.append(np.ascontiguousarray(b) if to is None else to(np.ascontiguousarray(b)))
        return MultiFab(boxes, arrs, ngrow, ncomp, nodal)

    def assemble(self, n_cell, zlo=0, origin=(0, 0, 0)):
        """the valid regions put back together: (ncomp, nz(+1), ny(+1), nx(+1))"""
        nx, ny, nz = n_cell
        ox, oy, oz = origin
        zlo = zlo + oz
        e = 1 if self.nodal else 0
        out = np.zeros((self.ncomp, nz + e, ny + e, nx + e))
        g = self.ngrow
        for (lo, hi), a in zip(self.boxes, self.arrays):
            a = a.detach().cpu().numpy() if hasattr(a, "detach") else a
            a = a.reshape((self.ncomp,) + tuple(a.shape)[-3:])
            v = a[:, g:a.shape[1] - g, g:a.shape[2] - g, g:a.shape[3] - g]
            out[:, lo[2] - zlo:hi[2] - zlo + 1 + e, lo[1] - oy:hi[1] - oy + 1 + e, lo[0] - ox:hi[0] - ox + 1 + e] = v
        return out


def _empty_like(like, shape):
    if isinstance(like, np.ndarray):
        return np.zeros(shape)
    import torch
    return torch.zeros(shape, dtype=torch.float64, device=like.device)


def nccl_unique_id():
    """the 128 bytes of an ncclUniqueId (call on rank 0, broadcast to the other ranks)"""
    buf = C.create_string_buffer(128)
    rc = _lib.lib().b200np_nccl_unique_id(buf)
    if rc != 0:
        raise ProjectionError(rc)
    return buf.raw


def slab_range(n_cell, bclo, rank, nranks):
    """(cell_lo, cell_hi, node_lo, node_hi) in z of `rank` (b200np_slab_range; needs no GPU)"""
    g = Geom()
    for d in range(3):
        g.n_cell[d] = int(n_cell[d]); g.dx[d] = 1.0; g.bc_lo[d] = int(bclo[d]); g.bc_hi[d] = int(bclo[d])
    v = [C.c_int() for _ in range(4)]
    rc = _lib.lib().b200np_slab_range(C.byref(g), rank, nranks, *[C.byref(x) for x in v])
    if rc != 0:
        raise ProjectionError(rc)
    return tuple(x.value for x in v)


class IncfloProjection:
    """Keeps one b200np handle alive across time steps (the reference rebuilds the
    projector every call, :181-193; caching the hierarchy gives identical results)."""

    def __init__(self, n_cell, dx, bclo, bchi, opts=None, device=0, rank=0, nranks=1, nccl_id=None):
        """n_cell is the GLOBAL domain; with nranks > 1 this rank owns the z slab slab_range(...)
        and every array passed to apply_nodal_projection is the rank's local box."""
        self._L = _lib.lib()
        self.n = tuple(int(x) for x in n_cell)
        g = Geom()
        for d in range(3):
            g.n_cell[d] = self.n[d]; g.dx[d] = float(dx[d]); g.bc_lo[d] = int(bclo[d]); g.bc_hi[d] = int(bchi[d])
        self.opts = opts if opts is not None else nodal_proj_opts()
        self.rank, self.nranks = int(rank), int(nranks)
        self.zlo = slab_range(self.n, bclo, rank, nranks)[0] if nranks > 1 else 0
        h = C.c_void_p()
        if nranks > 1:
            buf = C.create_string_buffer(bytes(nccl_id), 128)
            rc = self._L.b200np_create_dist(C.byref(h), C.byref(g), C.byref(self.opts), device, rank, nranks, buf)
        else:
            rc = self._L.b200np_create(C.byref(h), C.byref(g), C.byref(self.opts), device)
        if rc != 0:
            raise ProjectionError(rc)
        self._h = h
        self.stats = Stats()

    def apply_nodal_projection(self, velocity, ng, gp, p_nd, density=None, ngd=0, ro_0=1.0, velocity_o=None,
                               inflow_vel=None, scaling_factor=1.0, incremental=False, proj_for_small_dt=False,
                               mg_rtol=1e-11, mg_atol=1e-14):
        """incflo::ApplyNodalProjection(density, time, scaling_factor, incremental)."""
        z = self.zlo
        pv, bv, _ = _ptr_box(velocity, (-ng, -ng, z - ng), 3)
        po, _, _ = _ptr_box(velocity_o, (-ng, -ng, z - ng), 3)
        pr, br, _ = _ptr_box(density, (-ngd, -ngd, z - ngd), 1)
        pg, bg, _ = _ptr_box(gp, (0, 0, z), 3)
        pp, bp, _ = _ptr_box(p_nd, (0, 0, z), 1)
        pi, _, _ = _ptr_box(inflow_vel, (-ng, -ng, z - ng), 3)
        _lib.torch_sync()
        rc = self._L.b200np_apply_nodal_projection(self._h, pv, C.byref(bv), po, pr, C.byref(br) if br is not None else None,
                                                   float(ro_0), pg, C.byref(bg), pp, C.byref(bp), pi,
                                                   float(scaling_factor), int(incremental), int(proj_for_small_dt),
                                                   float(mg_rtol), float(mg_atol), C.byref(self.stats))
        if rc != 0:
            raise ProjectionError(rc)
        return self.stats

    def apply_nodal_projection_mf(self, velocity, gp, p_nd, density=None, ro_0=1.0, velocity_o=None, inflow_vel=None,
                                  scaling_factor=1.0, incremental=False, proj_for_small_dt=False, mg_rtol=1e-11, mg_atol=1e-14):
        """incflo::ApplyNodalProjection over multi-box MultiFabs (class MultiFab)"""
        r = lambda m: m.ref() if m is not None else None
        _lib.torch_sync()
        rc = self._L.b200np_apply_nodal_projection_mf(self._h, r(velocity), r(velocity_o), r(density), float(ro_0), r(gp), r(p_nd),
                                                      r(inflow_vel), float(scaling_factor), int(incremental), int(proj_for_small_dt),
                                                      float(mg_rtol), float(mg_atol), C.byref(self.stats))
        if rc != 0:
            raise ProjectionError(rc)
        return self.stats

    def project_mf(self, vel, sigma=None, const_sigma=1.0, phi=None, gphi=None, rtol=1e-11, atol=1e-14):
        """Hydro::NodalProjector::project + getPhi / getGradPhi over multi-box MultiFabs"""
        r = lambda m: m.ref() if m is not None else None
        _lib.torch_sync()
        rc = self._L.b200np_project_mf(self._h, r(vel), r(sigma), float(const_sigma), r(phi), r(gphi), float(rtol), float(atol),
                                       C.byref(self.stats))
        if rc != 0:
            raise ProjectionError(rc)
        return self.stats

    def set_inflow_profile(self, probtype, bcv_vel, time=0.0):
        """IncfloVelFill (src/prob/prob_bc.H) on the device: apply_nodal_projection(inflow_vel=None) then fills the
        first ghost layer at mass-inflow faces itself.  bcv_vel: 6 x 3 (Orientation x-lo, y-lo, z-lo, x-hi, y-hi, z-hi);
        None switches it off."""
        if bcv_vel is None:
            rc = self._L.b200np_set_inflow_profile(self._h, 0, None, 0.0)
        else:
            arr = (C.c_double * 18)(*[float(x) for x in np.asarray(bcv_vel, dtype=np.float64).reshape(18)])
            rc = self._L.b200np_set_inflow_profile(self._h, int(probtype), C.byref(arr), float(time))
        if rc != 0:
            raise ProjectionError(rc)

    def set_face_types(self, face_type, mixed_split_dir=0, mixed_half_num_cells=0):
        """incflo's face kinds beyond LinOpBCType (FACE_DEFAULT / FACE_DIRECTION_DEPENDENT / FACE_MIXED per Orientation
        x-lo, y-lo, z-lo, x-hi, y-hi, z-hi): direction_dependent faces get IncfloVelFill's copy-out branch and
        enforceInOutSolvability, mixed faces the overset mask of incflo::make_nodalBC_mask"""
        arr = (C.c_int * 6)(*[int(x) for x in face_type])
        rc = self._L.b200np_set_face_types(self._h, C.byref(arr), int(mixed_split_dir), int(mixed_half_num_cells))
        if rc != 0:
            raise ProjectionError(rc)

    def check_overset_mask(self, mask):
        """MLNodeLaplacian::setOversetMask argument check: True iff the int32 nodal mask (nz+1, ny+1, nx+1) -- numpy or
        torch cuda -- equals the mask set_face_types put in effect"""
        if hasattr(mask, "data_ptr"):
            ptr, shape = C.c_void_p(mask.data_ptr()), tuple(mask.shape)
        else:
            assert mask.dtype == np.int32 and mask.flags["C_CONTIGUOUS"]
            ptr, shape = C.c_void_p(mask.ctypes.data), mask.shape
        b = FabBox()
        for d in range(3):
            b.lo[d] = 0 if d < 2 else self.zlo
            b.hi[d] = b.lo[d] + shape[2 - d] - 1
        b.ncomp = 1
        rc = self._L.b200np_check_overset_mask(self._h, ptr, C.byref(b))
        if rc not in (0, 7):
            raise ProjectionError(rc)
        return rc == 0

    def inout_flux(self):
        """(influx, outflux) found by the last projection's enforceInOutSolvability"""
        a, b = C.c_double(), C.c_double()
        self._L.b200np_inout_flux(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def set_stream(self, cuda_stream):
        """run on the caller's stream (int handle of a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)"""
        self._L.b200np_set_stream(self._h, C.c_void_p(cuda_stream))

    def halo_transport(self):
        """0: one GPU; 1: NVLink peer memory (CUDA IPC); 2: ncclSend/ncclRecv fallback"""
        return self._L.b200np_halo_transport(self._h)

    def peer_map(self):
        """how the neighbours' arenas are mapped: 0 not mapped, 1 cuMem* allocation over a POSIX fd, 2 legacy CUDA IPC"""
        return self._L.b200np_peer_map(self._h)

    def time_op(self, lev, op, arg=1, reps=10):
        ms = C.c_double()
        rc = self._L.b200np_time_op(self._h, lev, op, arg, reps, C.byref(ms))
        if rc:
            raise ProjectionError(rc)
        return ms.value

    def level_dims(self, lev):
        n = (C.c_int * 3)(); nn = (C.c_int * 3)()
        self._L.b200np_level_dims(self._h, lev, C.byref(n), C.byref(nn))
        return tuple(n), tuple(nn)

    def close(self):
        if self._h is not None:
            self._L.b200np_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CompositeProjection:
    """incflo::ApplyNodalProjection / Hydro::NodalProjector with amr.max_level = 1 (BASELINE configs[3]):
    coarse level over the whole domain + ONE fine box at ratio 2, covered coarse cells [fine_lo, fine_hi]
    (inclusive).  Level-1 arrays live in FINE index space (cells 2*fine_lo .. 2*fine_hi+1), grown by
    their ghost width.  Backed by b200np_composite_* (include/b200np.h); algorithm: oracle/composite.py."""

    def __init__(self, n_cell, dx, bclo, bchi, fine_lo, fine_hi, opts=None, device=0):
        self._L = _lib.lib()
        self.n = tuple(int(x) for x in n_cell)
        self.flo = tuple(int(x) for x in fine_lo); self.fhi = tuple(int(x) for x in fine_hi)
        self.nf = tuple(2 * (h - l + 1) for l, h in zip(self.flo, self.fhi))
        g = Geom()
        for d in range(3):
            g.n_cell[d] = self.n[d]; g.dx[d] = float(dx[d]); g.bc_lo[d] = int(bclo[d]); g.bc_hi[d] = int(bchi[d])
        self.opts = opts if opts is not None else nodal_proj_opts()
        lo = (C.c_int * 3)(*self.flo); hi = (C.c_int * 3)(*self.fhi)
        h = C.c_void_p()
        rc = self._L.b200np_composite_create(C.byref(h), C.byref(g), C.byref(lo), C.byref(hi), C.byref(self.opts), device)
        if rc != 0:
            raise ProjectionError(rc)
        self._h = h
        self.stats = Stats()

    def _flo(self, ng):
        return tuple(2 * l - ng for l in self.flo)

    def project(self, vel0, ng0, vel1, ng1, sigma0=None, sigma1=None, const_sigma=1.0, rtol=1e-11, atol=1e-14,
                phi0=None, phi1=None, gphi0=None, gphi1=None):
        """Hydro::NodalProjector::project over both levels; vel0 / vel1 are updated in place.
        Returns (phi0, phi1, gphi0, gphi1) (allocated like vel0 when not given)."""
        n, nf = self.n, self.nf
        if phi0 is None:
            phi0 = _empty_like(vel0, (n[2] + 1, n[1] + 1, n[0] + 1))
        if phi1 is None:
            phi1 = _empty_like(vel1, (nf[2] + 1, nf[1] + 1, nf[0] + 1))
        if gphi0 is None:
            gphi0 = _empty_like(vel0, (3, n[2], n[1], n[0]))
        if gphi1 is None:
            gphi1 = _empty_like(vel1, (3, nf[2], nf[1], nf[0]))
        pv0, bv0, _ = _ptr_box(vel0, (-ng0,) * 3, 3)
        pv1, bv1, _ = _ptr_box(vel1, self._flo(ng1), 3)
        ps0, bs0, _ = _ptr_box(sigma0, (0, 0, 0), 1)
        ps1, bs1, _ = _ptr_box(sigma1, self._flo(0), 1)
        pp0, bp0, _ = _ptr_box(phi0, (0, 0, 0), 1)
        pp1, bp1, _ = _ptr_box(phi1, self._flo(0), 1)
        pg0, bg0, _ = _ptr_box(gphi0, (0, 0, 0), 3)
        pg1, bg1, _ = _ptr_box(gphi1, self._flo(0), 3)
        ref = lambda b: C.byref(b) if b is not None else None
        _lib.torch_sync()
        rc = self._L.b200np_composite_project(self._h, pv0, ref(bv0), pv1, ref(bv1), ps0, ref(bs0), ps1, ref(bs1),
                                              float(const_sigma), pp0, ref(bp0), pp1, ref(bp1), pg0, ref(bg0), pg1, ref(bg1),
                                              float(rtol), float(atol), C.byref(self.stats))
        if rc != 0:
            raise ProjectionError(rc)
        return phi0, phi1, gphi0, gphi1

    def apply_nodal_projection(self, velocity, ng, gp, p_nd, density=None, ngd=(0, 0), ro_0=1.0, velocity_o=None,
                               inflow_vel=None, scaling_factor=1.0, incremental=False, proj_for_small_dt=False,
                               mg_rtol=1e-11, mg_atol=1e-14):
        """incflo::ApplyNodalProjection with finest_level = 1: velocity, gp, p_nd, density, velocity_o are
        pairs (level 0, level 1); ng / ngd the ghost widths per level."""
        fbp = C.POINTER(FabBox)

        def pair(arrs, los, ncomp):
            ptrs = (C.c_void_p * 2)(); boxes = (fbp * 2)(); keep = []
            if arrs is None:
                return None, None, keep
            for l in range(2):
                p, b, _ = _ptr_box(arrs[l], los[l], ncomp)
                ptrs[l] = p
                if b is not None:
                    keep.append(b); boxes[l] = C.pointer(b)
            return ptrs, boxes, keep
        lo_v = ((-ng[0],) * 3, self._flo(ng[1]))
        lo_d = ((-ngd[0],) * 3, self._flo(ngd[1]))
        lo_0 = ((0, 0, 0), self._flo(0))
        pv, bv, k1 = pair(velocity, lo_v, 3)
        po, _, k2 = pair(velocity_o, lo_v, 3)
        pr, br, k3 = pair(density, lo_d, 1)
        pg, bg, k4 = pair(gp, lo_0, 3)
        pp, bp, k5 = pair(p_nd, lo_0, 1)
        pi, _, _ = _ptr_box(inflow_vel, (-ng[0],) * 3, 3)
        ref = lambda a: C.byref(a) if a is not None else None
        _lib.torch_sync()
        rc = self._L.b200np_composite_apply_nodal_projection(self._h, ref(pv), ref(bv), ref(po), ref(pr), ref(br), float(ro_0),
                                                             ref(pg), ref(bg), ref(pp), ref(bp), pi, float(scaling_factor),
                                                             int(incremental), int(proj_for_small_dt), float(mg_rtol),
                                                             float(mg_atol), C.byref(self.stats))
        if rc != 0:
            raise ProjectionError(rc)
        return self.stats

    def apply_nodal_projection_mf(self, velocity, gp, p_nd, density=None, ro_0=1.0, velocity_o=None, inflow_vel=None,
                                  scaling_factor=1.0, incremental=False, proj_for_small_dt=False, mg_rtol=1e-11, mg_atol=1e-14):
        """the same over multi-box MultiFabs: velocity, gp, p_nd, density, velocity_o are pairs (level 0, level 1) of class
        MultiFab; the boxes of level 1 tile the fine box in fine index space (MultiFab.split(..., origin=2 * fine_lo))"""
        mfp = C.POINTER(MFab)

        def pair(ms):
            if ms is None:
                return None
            a = (mfp * 2)()
            for l in range(2):
                a[l] = C.pointer(ms[l].c)
            return a
        keep = [pair(velocity), pair(velocity_o), pair(density), pair(gp), pair(p_nd)]
        ref = lambda a: C.byref(a) if a is not None else None
        _lib.torch_sync()
        rc = self._L.b200np_composite_apply_nodal_projection_mf(self._h, ref(keep[0]), ref(keep[1]), ref(keep[2]), float(ro_0), ref(keep[3]),
                                                                ref(keep[4]), inflow_vel.ref() if inflow_vel is not None else None,
                                                                float(scaling_factor), int(incremental), int(proj_for_small_dt),
                                                                float(mg_rtol), float(mg_atol), C.byref(self.stats))
        if rc != 0:
            raise ProjectionError(rc)
        return self.stats

    def set_stream(self, cuda_stream):
        self._L.b200np_composite_set_stream(self._h, C.c_void_p(cuda_stream))

    def close(self):
        if self._h is not None:
            self._L.b200np_composite_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
