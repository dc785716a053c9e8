"""Loader + build recipe of the C-ABI library libb200np.so (include/b200np.h).

The library is hand-written sm_100a CUDA behind a plain C ABI; this module only
dlopens it.  There is no CPU fallback: if the library is missing or no CUDA
device is usable the product path raises.
"""
import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIBDIR = os.path.join(_PKG, "lib")
SO = os.path.join(LIBDIR, "libb200np.so")
SOURCES = ["b200np.cu", "b200mac.cu", "b200eb.cu"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    # every file of csrc/ (sources and headers alike) and the public header
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "b200np.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> incflo_b200/lib/libb200np.so"""
    if not force and not _stale():
        return SO
    os.makedirs(LIBDIR, exist_ok=True)
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    nvcc = os.environ.get("NVCC", os.path.join(cuda_home, "bin", "nvcc"))
    ccbin = os.environ.get("B200NP_CCBIN", "/usr/bin/g++")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    # one nvcc per translation unit, in parallel (the three files share no device symbols), then one link
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + ["-ccbin", ccbin]
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append(subprocess.Popen([nvcc] + flags + ["-c", "-o", obj, os.path.join(CSRC, src)], env=env))
    rcs = [p.wait() for p in procs]
    if any(rcs):
        raise subprocess.CalledProcessError(max(rcs), "nvcc -c " + " ".join(SOURCES))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-ccbin", ccbin, "-o", SO] + objs, env=env)
    return SO


class Geom(C.Structure):
    _fields_ = [("n_cell", C.c_int * 3), ("dx", C.c_double * 3), ("bc_lo", C.c_int * 3), ("bc_hi", C.c_int * 3)]


class Opts(C.Structure):
    _fields_ = [("verbose", C.c_int), ("bottom_verbose", C.c_int), ("maxiter", C.c_int), ("bottom_maxiter", C.c_int),
                ("bottom_rtol", C.c_double), ("bottom_atol", C.c_double), ("mg_max_coarsening_level", C.c_int),
                ("num_pre_smooth", C.c_int), ("num_post_smooth", C.c_int), ("smooth_num_sweeps", C.c_int),
                ("bottom_solver", C.c_int), ("tile", C.c_int * 3), ("use_graph", C.c_int)]


class FabBox(C.Structure):
    _fields_ = [("lo", C.c_int * 3), ("hi", C.c_int * 3), ("ncomp", C.c_int)]


class MFab(C.Structure):
    """b200np_mfab: the local FArrayBoxes of an amrex::MultiFab"""
    _fields_ = [("nfabs", C.c_int), ("ngrow", C.c_int), ("ncomp", C.c_int), ("box", C.POINTER(FabBox)), ("data", C.POINTER(C.c_void_p))]


class EBFlow(C.Structure):
    """b200eb_flow: the eb_flow.* inputs of set_eb_velocity / density / tracer"""
    _fields_ = [("has_normal", C.c_int), ("normal", C.c_double * 3), ("normal_tol", C.c_double), ("is_mag", C.c_int), ("vel_mag", C.c_double),
                ("velocity", C.c_double * 3), ("density", C.c_double), ("ntrac", C.c_int), ("tracer", C.c_double * 8)]


class Stats(C.Structure):
    _fields_ = [("iters", C.c_int), ("nlevels", C.c_int), ("bottom_iters", C.c_int), ("status", C.c_int),
                ("rhsnorm", C.c_double), ("resnorm0", C.c_double), ("resnorm", C.c_double),
                ("resnorm_hist", C.c_double * 128), ("ms_total", C.c_double), ("ms_solve", C.c_double),
                ("ms_h2d", C.c_double), ("ms_d2h", C.c_double), ("h2d_bytes", C.c_longlong),
                ("d2h_bytes", C.c_longlong), ("launches", C.c_longlong)]


# every symbol include/b200np.h declares
EXPORTS = ["b200np_default_opts", "b200np_create", "b200np_create_dist", "b200np_nccl_unique_id", "b200np_slab_range", "b200np_dist_plan",
           "b200np_destroy", "b200np_set_stream", "b200np_project",
           "b200np_apply_nodal_projection", "b200np_project_mf", "b200np_apply_nodal_projection_mf", "b200np_set_inflow_profile", "b200np_set_face_types", "b200np_check_overset_mask", "b200np_inout_flux", "b200np_strerror", "b200np_version", "b200np_nlevels",
           "b200np_level_dims", "b200np_halo_transport", "b200np_peer_map", "b200np_set_sigma", "b200np_level_set", "b200np_level_get", "b200np_level_op",
           "b200np_time_op", "b200np_composite_create", "b200np_composite_destroy", "b200np_composite_set_stream",
           "b200np_composite_level", "b200np_composite_project", "b200np_composite_apply_nodal_projection", "b200np_composite_apply_nodal_projection_mf",
           "b200mac_create", "b200mac_destroy", "b200mac_nlevels", "b200mac_set_coeffs", "b200mac_project", "b200mac_level_op", "b200mac_level_dims",
           "b200mac_set_coeffs_mf", "b200mac_project_mf", "b200mac_set_stream", "b200eb_set_stream",
           "b200eb_create", "b200eb_destroy", "b200eb_nlevels", "b200eb_set_geometry", "b200eb_set_eb_inflow_velocity", "b200eb_set_eb_flow",
           "b200eb_project", "b200eb_apply_nodal_projection", "b200eb_build_stencils", "b200eb_level_stencil", "b200eb_level_op", "b200eb_level_dims",
           "b200eb_compute_rhs", "b200eb_time_op", "b200eb_set_geometry_mf", "b200eb_project_mf", "b200eb_apply_nodal_projection_mf"]

_lib = None


def torch_sync():
    """Device tensors handed to the library must be complete when the call is made: a handle runs on its own CUDA stream unless
    *_set_stream gave it the caller's.  If torch is in use, wait for its current stream (a no-op when that stream is idle)."""
    import sys
    t = sys.modules.get("torch")
    if t is not None and t.cuda.is_available() and t.cuda.is_initialized():
        t.cuda.current_stream().synchronize()


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise RuntimeError("libb200np.so is not built (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(SO)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.POINTER(C.c_int * 3)
    fb = C.POINTER(FabBox)
    L.b200np_default_opts.argtypes = [C.POINTER(Opts)]
    L.b200np_default_opts.restype = None
    L.b200np_create.argtypes = [C.POINTER(vp), C.POINTER(Geom), C.POINTER(Opts), C.c_int]
    L.b200np_create_dist.argtypes = [C.POINTER(vp), C.POINTER(Geom), C.POINTER(Opts), C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.b200np_nccl_unique_id.argtypes = [C.c_void_p]
    L.b200np_slab_range.argtypes = [C.POINTER(Geom), C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 4
    L.b200np_dist_plan.argtypes = [C.POINTER(Geom), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.b200np_destroy.argtypes = [vp]
    L.b200np_destroy.restype = None
    L.b200np_set_stream.argtypes = [vp, C.c_void_p]
    L.b200np_project.argtypes = [vp, dp, fb, dp, fb, C.c_double, dp, fb, dp, fb, C.c_double, C.c_double, C.POINTER(Stats)]
    L.b200np_apply_nodal_projection.argtypes = [vp, dp, fb, dp, dp, fb, C.c_double, dp, fb, dp, fb, dp, C.c_double,
                                                C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(Stats)]
    mf = C.POINTER(MFab)
    L.b200np_project_mf.argtypes = [vp, mf, mf, C.c_double, mf, mf, C.c_double, C.c_double, C.POINTER(Stats)]
    L.b200np_apply_nodal_projection_mf.argtypes = [vp, mf, mf, mf, C.c_double, mf, mf, mf, C.c_double, C.c_int, C.c_int, C.c_double,
                                                   C.c_double, C.POINTER(Stats)]
    L.b200np_set_inflow_profile.argtypes = [vp, C.c_int, C.POINTER(C.c_double * 18), C.c_double]
    L.b200np_set_face_types.argtypes = [vp, C.POINTER(C.c_int * 6), C.c_int, C.c_int]
    L.b200np_check_overset_mask.argtypes = [vp, C.c_void_p, fb]
    L.b200np_inout_flux.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.b200np_strerror.argtypes = [C.c_int]
    L.b200np_strerror.restype = C.c_char_p
    L.b200np_version.restype = C.c_int
    L.b200np_nlevels.argtypes = [vp]
    L.b200np_halo_transport.argtypes = [vp]
    L.b200np_peer_map.argtypes = [vp]
    L.b200np_level_dims.argtypes = [vp, C.c_int, ip, ip]
    L.b200np_set_sigma.argtypes = [vp, dp, fb, C.c_double]
    L.b200np_level_set.argtypes = [vp, C.c_int, C.c_int, dp]
    L.b200np_level_get.argtypes = [vp, C.c_int, C.c_int, dp]
    L.b200np_level_op.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.b200np_time_op.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    i3 = C.POINTER(C.c_int * 3)
    L.b200np_composite_create.argtypes = [C.POINTER(vp), C.POINTER(Geom), i3, i3, C.POINTER(Opts), C.c_int]
    L.b200np_composite_destroy.argtypes = [vp]
    L.b200np_composite_destroy.restype = None
    L.b200np_composite_set_stream.argtypes = [vp, C.c_void_p]
    L.b200np_composite_level.argtypes = [vp, C.c_int]
    L.b200np_composite_level.restype = vp
    L.b200np_composite_project.argtypes = [vp, dp, fb, dp, fb, dp, fb, dp, fb, C.c_double, dp, fb, dp, fb, dp, fb, dp, fb,
                                           C.c_double, C.c_double, C.POINTER(Stats)]
    p2, f2 = C.POINTER(C.c_void_p * 2), C.POINTER(fb * 2)
    L.b200np_composite_apply_nodal_projection.argtypes = [vp, p2, f2, p2, p2, f2, C.c_double, p2, f2, p2, f2, dp, C.c_double,
                                                          C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(Stats)]
    m2 = C.POINTER(mf * 2)
    L.b200np_composite_apply_nodal_projection_mf.argtypes = [vp, m2, m2, m2, C.c_double, m2, m2, mf, C.c_double, C.c_int, C.c_int,
                                                             C.c_double, C.c_double, C.POINTER(Stats)]
    L.b200mac_create.argtypes = [C.POINTER(vp), C.POINTER(Geom), C.POINTER(Opts), C.c_int]
    L.b200mac_destroy.argtypes = [vp]
    L.b200mac_destroy.restype = None
    L.b200mac_nlevels.argtypes = [vp]
    L.b200mac_set_coeffs.argtypes = [vp, dp, fb, dp, fb, dp, fb, C.c_double]
    L.b200mac_project.argtypes = [vp, dp, fb, dp, fb, dp, fb, dp, fb, C.c_int, C.c_double, C.c_double, C.POINTER(Stats)]
    L.b200mac_level_op.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp, dp]
    L.b200mac_level_dims.argtypes = [vp, C.c_int, ip]
    L.b200mac_set_coeffs_mf.argtypes = [vp, mf, mf, mf]
    L.b200mac_project_mf.argtypes = [vp, mf, mf, mf, mf, C.c_int, C.c_double, C.c_double, C.POINTER(Stats)]
    L.b200eb_create.argtypes = [C.POINTER(vp), C.POINTER(Geom), C.POINTER(Opts), C.c_int]
    L.b200eb_destroy.argtypes = [vp]
    L.b200eb_destroy.restype = None
    L.b200eb_nlevels.argtypes = [vp]
    L.b200eb_set_geometry.argtypes = [vp, dp, fb, dp, fb]
    L.b200eb_set_eb_inflow_velocity.argtypes = [vp, dp, fb, dp, fb, dp, fb]
    L.b200eb_set_eb_flow.argtypes = [vp, C.POINTER(EBFlow), C.c_int, dp, fb, dp, fb, dp, fb, dp, fb]
    L.b200eb_project.argtypes = [vp, dp, fb, dp, fb, C.c_double, dp, fb, dp, fb, C.c_double, C.c_double, C.POINTER(Stats)]
    L.b200eb_apply_nodal_projection.argtypes = [vp, dp, fb, dp, dp, fb, C.c_double, dp, fb, dp, fb, dp, C.c_double,
                                                C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(Stats)]
    L.b200eb_build_stencils.argtypes = [vp, dp, fb, C.c_double]
    L.b200eb_level_stencil.argtypes = [vp, C.c_int, dp]
    L.b200eb_level_op.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp, dp]
    L.b200eb_level_dims.argtypes = [vp, C.c_int, ip, ip]
    L.b200eb_compute_rhs.argtypes = [vp, dp, fb, dp]
    L.b200eb_set_stream.argtypes = [vp, C.c_void_p]
    L.b200mac_set_stream.argtypes = [vp, C.c_void_p]
    L.b200eb_set_geometry_mf.argtypes = [vp, mf, mf]
    L.b200eb_project_mf.argtypes = [vp, mf, mf, C.c_double, mf, mf, C.c_double, C.c_double, C.POINTER(Stats)]
    L.b200eb_apply_nodal_projection_mf.argtypes = [vp, mf, mf, mf, C.c_double, mf, mf, mf, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                                   C.POINTER(Stats)]
    L.b200eb_time_op.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    _lib = L
    return L
