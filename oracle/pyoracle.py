"""ctypes binding of the CPU oracle (oracle/nodal_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(incflo_b200) never imports this module.  PARITY UNPINNED -- see the header of
nodal_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BC_PERIODIC, BC_NEUMANN, BC_DIRICHLET, BC_INFLOW = 0, 1, 2, 3
SM_LEX, SM_COLOR8, SM_COLOR4XY, SM_JACOBI, SM_BOX, SM_PLANE4 = 0, 1, 2, 3, 4, 5


class Params(C.Structure):
    _fields_ = [
        ("n", C.c_int * 3), ("dx", C.c_double * 3), ("bclo", C.c_int * 3), ("bchi", C.c_int * 3),
        ("max_coarsening_level", C.c_int), ("maxiter", C.c_int), ("bottom_maxiter", C.c_int),
        ("bottom_rtol", C.c_double), ("bottom_atol", C.c_double),
        ("nu1", C.c_int), ("nu2", C.c_int), ("nsweeps", C.c_int), ("smoother", C.c_int),
        ("box", C.c_int * 3), ("box_order", C.c_int), ("box_stale_per_call", C.c_int), ("verbose", C.c_int),
        ("box_amrex", C.c_int),
        ("mixed_lo", C.c_int * 3), ("mixed_hi", C.c_int * 3), ("mix_dir", C.c_int), ("mix_half", C.c_int),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("iters", C.c_int), ("nlevels", C.c_int), ("bottom_iters", C.c_int), ("status", C.c_int),
        ("rhsnorm", C.c_double), ("resnorm0", C.c_double), ("resnorm", C.c_double),
        ("resnorm_hist", C.c_double * 128), ("t_solve", C.c_double), ("t_total", C.c_double),
    ]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "nodal_oracle.c")
    if force or not os.path.exists(so) or (
            os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        L = _LIB
        L.orc_default_params.argtypes = [C.POINTER(Params)]
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_set_num_threads.restype = C.c_int
        L.orc_project.argtypes = [C.POINTER(Params), dp, C.c_int, dp, C.c_double, dp, dp, dp, C.c_double,
                                  C.c_double, C.POINTER(Stats)]
        L.orc_project.restype = C.c_int
        L.orc_apply_nodal_projection.argtypes = [C.POINTER(Params), dp, dp, C.c_int, dp, C.c_int, C.c_double, dp, dp,
                                                 dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                                 C.POINTER(Stats)]
        L.orc_apply_nodal_projection.restype = C.c_int
        L.orc_mg_create.argtypes = [C.POINTER(Params), dp, C.c_double]
        L.orc_mg_create.restype = C.c_void_p
        L.orc_mg_destroy.argtypes = [C.c_void_p]
        L.orc_mg_nlevels.argtypes = [C.c_void_p]
        L.orc_mg_nlevels.restype = C.c_int
        L.orc_mg_level_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int * 3), C.POINTER(C.c_int * 3)]
        L.orc_mg_sigma.argtypes = [C.c_void_p, C.c_int]
        L.orc_mg_sigma.restype = dp
        L.orc_adotx.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.orc_residual.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
        L.orc_smooth.argtypes = [C.c_void_p, C.c_int, dp, dp, C.c_int]
        L.orc_restrict.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.orc_interp_add.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.orc_divu.argtypes = [C.c_void_p, dp, C.c_int, dp]
        L.orc_mknewu.argtypes = [C.c_void_p, dp, dp, C.c_int, dp]
        L.orc_bottom_solve.argtypes = [C.c_void_p, dp, dp]
        L.orc_bottom_solve.restype = C.c_int
        L.orc_mlmg_solve.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_double, C.POINTER(Stats)]
        L.orc_mlmg_solve.restype = C.c_int
        L.orc_dot_weight.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_dot_weight.restype = C.c_double
        L.orc_dot_weights.argtypes = [C.c_void_p, C.c_int, dp]
    return _LIB


def set_num_threads(n):
    """OpenMP threads of the oracle from now on (torchrun exports OMP_NUM_THREADS=1); returns the count in effect"""
    return lib().orc_set_num_threads(int(n))


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_params(n, dx, bclo=(0, 0, 0), bchi=(0, 0, 0), **kw):
    p = Params()
    lib().orc_default_params(C.byref(p))
    for d in range(3):
        p.n[d] = int(n[d]); p.dx[d] = float(dx[d]); p.bclo[d] = int(bclo[d]); p.bchi[d] = int(bchi[d])
    for k, v in kw.items():
        if k in ("box", "mixed_lo", "mixed_hi"):
            for d in range(3):
                getattr(p, k)[d] = int(v[d])
        else:
            setattr(p, k, v)
    return p


class MG:
    """Multigrid hierarchy handle for per-kernel parity tests (unique-node layout).

    Arrays are numpy float64 of shape (nnz, nny, nnx) (C order == i fastest)."""

    def __init__(self, params, sigma=None, const_sigma=1.0):
        self.params = params
        self.h = lib().orc_mg_create(C.byref(params), _p(sigma), float(const_sigma))
        self.nlev = lib().orc_mg_nlevels(self.h)

    def close(self):
        if self.h:
            lib().orc_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def dims(self, lev):
        n = (C.c_int * 3)(); nn = (C.c_int * 3)()
        lib().orc_mg_level_dims(self.h, lev, C.byref(n), C.byref(nn))
        return tuple(n), tuple(nn)

    def node_shape(self, lev):
        _, nn = self.dims(lev)
        return (nn[2], nn[1], nn[0])

    def sigma(self, lev):
        n, _ = self.dims(lev)
        ptr = lib().orc_mg_sigma(self.h, lev)
        if not ptr:
            return None
        return np.ctypeslib.as_array(ptr, shape=(n[2], n[1], n[0])).copy()

    def adotx(self, lev, phi):
        y = np.empty_like(phi); lib().orc_adotx(self.h, lev, _p(phi), _p(y)); return y

    def residual(self, lev, phi, rhs):
        r = np.empty_like(phi); lib().orc_residual(self.h, lev, _p(phi), _p(rhs), _p(r)); return r

    def smooth(self, lev, phi, rhs, nsweeps):
        lib().orc_smooth(self.h, lev, _p(phi), _p(rhs), nsweeps); return phi

    def restrict(self, flev, fine):
        c = np.empty(self.node_shape(flev + 1)); lib().orc_restrict(self.h, flev, _p(fine), _p(c)); return c

    def interp_add(self, flev, fine, crse):
        lib().orc_interp_add(self.h, flev, _p(fine), _p(crse)); return fine

    def divu(self, vel, ng):
        r = np.empty(self.node_shape(0)); lib().orc_divu(self.h, _p(vel), ng, _p(r)); return r

    def mknewu(self, phi, vel, ng):
        n, _ = self.dims(0)
        g = np.empty((3, n[2], n[1], n[0])); lib().orc_mknewu(self.h, _p(phi), _p(vel), ng, _p(g)); return g

    def bottom_solve(self, b):
        x = np.zeros_like(b); it = lib().orc_bottom_solve(self.h, _p(x), _p(b)); return x, it

    def solve(self, phi, rhs, rtol=1e-11, atol=1e-14):
        st = Stats(); lib().orc_mlmg_solve(self.h, _p(phi), _p(rhs), rtol, atol, C.byref(st)); return st

    def dot_weights(self, lev):
        _, nn = self.dims(lev)
        w = np.empty((nn[2], nn[1], nn[0]))
        lib().orc_dot_weights(self.h, lev, _p(w))
        return w


def project(params, vel, ng, sigma=None, const_sigma=1.0, rtol=1e-11, atol=1e-14, want_rhs=False):
    """Hydro::NodalProjector::project.  vel: (3, nz+2ng, ny+2ng, nx+2ng) in/out."""
    n = tuple(params.n)
    phi = np.zeros((n[2] + 1, n[1] + 1, n[0] + 1))
    gphi = np.zeros((3, n[2], n[1], n[0]))
    rhs = np.zeros_like(phi) if want_rhs else None
    st = Stats()
    status = lib().orc_project(C.byref(params), _p(vel), ng, _p(sigma), float(const_sigma), _p(phi), _p(gphi),
                               _p(rhs), rtol, atol, C.byref(st))
    return dict(status=status, phi=phi, gphi=gphi, rhs=rhs, stats=st)


def apply_nodal_projection(params, velocity, ng, gp, p_nd, density=None, ngd=0, ro_0=1.0, velocity_o=None,
                           inflow_vel=None, scaling_factor=1.0, incremental=False, proj_for_small_dt=False,
                           rtol=1e-11, atol=1e-14):
    st = Stats()
    status = lib().orc_apply_nodal_projection(C.byref(params), _p(velocity), _p(velocity_o), ng, _p(density), ngd,
                                              float(ro_0), _p(gp), _p(p_nd), _p(inflow_vel), float(scaling_factor),
                                              int(incremental), int(proj_for_small_dt), rtol, atol, C.byref(st))
    return status, st
