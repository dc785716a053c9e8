"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/b200np.h declares, mirrors the reference's defaults and error behaviour, and fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu


def _lib():
    from incflo_b200 import _lib
    return _lib, _lib.lib()


def test_library_exports_every_declared_symbol():
    mod, L = _lib()
    header = open(os.path.join(ROOT, "include", "b200np.h")).read()
    declared = set(re.findall(r"\b(b200(?:np|mac|eb)_[a-z_]+)\s*\(", header))
    assert declared == set(mod.EXPORTS), declared ^ set(mod.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.b200np_version() == 1


def test_default_opts_match_reference_defaults():
    mod, L = _lib()
    o = mod.Opts()
    L.b200np_default_opts(C.byref(o))
    # src/incflo.H:449-458 and Docs/sphinx_documentation/source/InputsMultigrid.rst:10-36
    assert (o.verbose, o.bottom_verbose, o.maxiter, o.bottom_maxiter) == (0, 0, 100, 100)
    assert o.bottom_rtol == 1e-4 and o.bottom_atol == -1.0
    assert o.mg_max_coarsening_level == 100
    assert (o.num_pre_smooth, o.num_post_smooth, o.smooth_num_sweeps) == (2, 2, 4)


def test_projection_bc_mapping():
    """incflo::get_projection_bc (src/projection/incflo_projection_bc.cpp:5-41)"""
    from incflo_b200 import nodal_projector as npj
    assert npj.get_projection_bc((1, 0, 0), ("x", "po", "mi")) == (npj.BC_PERIODIC, npj.BC_DIRICHLET, npj.BC_INFLOW)
    assert npj.get_projection_bc((0, 0, 0), ("pi", "sw", "nsw")) == (npj.BC_DIRICHLET, npj.BC_NEUMANN, npj.BC_NEUMANN)
    assert npj.get_projection_bc((0, 0, 0), ("dd", "mixed", "mass_inflow")) == (npj.BC_INFLOW,) * 3
    with pytest.raises(npj.ProjectionError):  # "get_projection_bc: undefined BC type" aborts in the reference
        npj.get_projection_bc((0, 0, 0), ("bogus", "sw", "sw"))


def test_bad_arguments_are_rejected():
    mod, L = _lib()
    g = mod.Geom()
    h = C.c_void_p()
    for d in range(3):
        g.n_cell[d] = 8; g.dx[d] = 0.1
    g.bc_lo[0] = 7
    assert L.b200np_create(C.byref(h), C.byref(g), None, 0) == 3  # B200NP_ERR_BAD_BC
    g.bc_lo[0] = 0; g.bc_hi[0] = 1  # periodic on one side only
    assert L.b200np_create(C.byref(h), C.byref(g), None, 0) == 3
    g.bc_hi[0] = 0; g.n_cell[1] = 0
    assert L.b200np_create(C.byref(h), C.byref(g), None, 0) == 4  # B200NP_ERR_BAD_ARG
    assert b"undefined BC type" in L.b200np_strerror(3)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    mod, L = _lib()
    g = mod.Geom()
    for d in range(3):
        g.n_cell[d] = 8; g.dx[d] = 0.1
    h = C.c_void_p()
    rc = L.b200np_create(C.byref(h), C.byref(g), None, 0)
    assert rc == 5 and not h.value  # B200NP_ERR_CUDA: fails loudly, nothing computed on the CPU
    from incflo_b200 import nodal_projector as npj
    import numpy as np
    proj = npj.NodalProjector(np.zeros((3, 10, 10, 10)), None, 1.0, dict(n_cell=(8, 8, 8), dx=(0.1,) * 3), ng=1)
    with pytest.raises(npj.ProjectionError):
        proj.project(1e-11, 1e-14)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under incflo_b200/ may reference it"""
    pkg = os.path.join(ROOT, "incflo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in src and "nodal_oracle" not in src and "liboracle" not in src, f


def test_composite_create_validates_the_fine_box_without_a_gpu():
    """b200np_composite_create: which fine boxes are supported is decided before any CUDA call"""
    mod, L = _lib()
    g = mod.Geom()
    for d in range(3):
        g.n_cell[d] = 16; g.dx[d] = 1 / 16
    g.bc_lo[0] = g.bc_hi[0] = 0          # periodic x
    g.bc_lo[1] = 3; g.bc_hi[1] = 2       # inflow y-lo, outflow y-hi
    g.bc_lo[2] = g.bc_hi[2] = 1          # walls z

    def create(lo, hi):
        h = C.c_void_p()
        rc = L.b200np_composite_create(C.byref(h), C.byref(g), C.byref((C.c_int * 3)(*lo)), C.byref((C.c_int * 3)(*hi)), None, 0)
        if rc == 0:
            L.b200np_composite_destroy(h)
        return rc
    ok = 0 if has_gpu() else 5                                  # a supported box needs a device (no CPU fallback)
    assert create((4, 4, 4), (11, 11, 11)) == ok               # interior box
    assert create((0, 4, 4), (15, 11, 11)) == ok               # spans the periodic direction
    assert create((4, 4, 0), (11, 11, 7)) == ok                # touches a wall
    assert create((4, 8, 4), (11, 15, 11)) == ok               # touches the outflow face
    assert create((0, 4, 4), (7, 11, 11)) == 7                 # on the periodic seam without spanning: unsupported
    assert create((4, 0, 4), (11, 7, 11)) == 7                 # on the inflow face: unsupported
    assert create((4, 4, 4), (3, 11, 11)) == 4                 # hi < lo
    assert create((4, 4, 4), (11, 11, 16)) == 4                # outside the domain
    g.bc_lo[1] = g.bc_hi[1] = 1
    assert create((0, 0, 0), (15, 15, 15)) == 7                # the "fine box" is the whole domain


def test_inflow_profile_argument_checks():
    mod, L = _lib()
    arr = (C.c_double * 18)()
    assert L.b200np_set_inflow_profile(None, 31, C.byref(arr), 0.0) == 4   # no handle


def test_eb_argument_checks_need_no_gpu():
    """b200eb_create validates geometry and BCs before touching CUDA (AMReX's EB support asserts dx == dy == dz)"""
    mod, L = _lib()
    g = mod.Geom()
    for d in range(3):
        g.n_cell[d] = 8; g.dx[d] = 0.125; g.bc_lo[d] = 0; g.bc_hi[d] = 0
    h = C.c_void_p()
    g.dx[2] = 0.25
    assert L.b200eb_create(C.byref(h), C.byref(g), None, 0) == 4       # B200NP_ERR_BAD_ARG
    g.dx[2] = 0.125
    g.bc_lo[0] = 1                                                      # periodic on one side only
    assert L.b200eb_create(C.byref(h), C.byref(g), None, 0) == 3       # B200NP_ERR_BAD_BC
    g.bc_lo[0] = 0
    if not has_gpu():
        assert L.b200eb_create(C.byref(h), C.byref(g), None, 0) == 5   # B200NP_ERR_CUDA: no CPU fallback
    assert L.b200eb_project(None, None, None, None, None, 1.0, None, None, None, None, 1e-11, 1e-14, None) == 4
    assert L.b200eb_nlevels(None) == 0


def test_multifab_split_and_assemble_round_trip_with_an_origin():
    """host logic of the multi-box mirror: chopping a single-box array into boxes (own ghost frames, optional index origin as a fine AMR
    level has it) and putting the valid regions back together is the identity; the b200np_mfab view carries the allocated boxes"""
    import numpy as np
    from incflo_b200 import nodal_projector as npj
    rng = np.random.default_rng(5)
    n, ng, org = (10, 6, 8), 2, (16, 8, 12)
    full = rng.standard_normal((3, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng))
    m = npj.MultiFab.split(full, n, 4, ng, 3, origin=org)
    assert len(m.boxes) == 3 * 2 * 2 and m.c.nfabs == 12 and m.c.ngrow == ng and m.c.ncomp == 3
    assert m.boxes[0] == ((16, 8, 12), (19, 11, 15))
    assert m.boxes[-1] == ((24, 12, 16), (25, 13, 19))           # ragged last boxes: 2 x 2 x 4 cells
    b0 = m.c.box[0]
    assert [b0.lo[d] for d in range(3)] == [14, 6, 10] and [b0.hi[d] for d in range(3)] == [21, 13, 17]   # valid box grown by ngrow
    assert np.array_equal(m.assemble(n, origin=org), full[:, ng:-ng, ng:-ng, ng:-ng])
    # interior ghost cells of a box hold the neighbour's valid values
    lo, hi = m.boxes[1]
    a = m.arrays[1]
    assert np.array_equal(a[:, ng:-ng, ng:-ng, 0:ng], full[:, ng:ng + 4, ng:ng + 4, ng + 4 - ng:ng + 4])
    pn = rng.standard_normal((1, n[2] + 1, n[1] + 1, n[0] + 1))
    mn = npj.MultiFab.split(pn, n, 4, 0, 1, nodal=True, origin=org)
    assert np.array_equal(mn.assemble(n, origin=org), pn)
    q = mn.c.box[0]
    assert [q.hi[d] - q.lo[d] + 1 for d in range(3)] == [5, 5, 5]   # nodal: one more node than cells per direction
