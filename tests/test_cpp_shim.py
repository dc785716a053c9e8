"""The C++ host-side mirror of the reference interface (include/B200NodalProjector.H) over the C ABI:
host logic on CPU, and -- on the GPU box -- the call sequence of
incflo_apply_nodal_projection.cpp:181-219 compiled as plain C++ against libb200np.so, checked
against the CPU oracle (1e-9 relative L2, north_star tolerance)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import oracle_params, rel_l2, remove_mean

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim_exe(tmp_path_factory):
    from incflo_b200 import _lib
    _lib.build()
    exe = str(tmp_path_factory.mktemp("shim") / "shim_check")
    libdir = os.path.dirname(_lib.SO)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "shim_check.cpp"), "-L" + libdir, "-lb200np",
                           "-Wl,-rpath," + libdir])
    return exe


def test_shim_host_logic(shim_exe):
    out = subprocess.run([shim_exe, "host"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shim host checks OK" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("config,N", [("tgv", 32), ("rt", 32)])
def test_shim_project_matches_oracle(shim_exe, tmp_path, config, N, oracle):
    from incflo_b200 import problems
    cfg = problems.make(config, N, ng=1, device="cpu")
    vel = cfg["vel"].numpy().copy()
    sigma = None if cfg["sigma"] is None else cfg["sigma"].numpy().copy()
    if sigma is None:
        cfg["const_sigma"] = 0.37  # the value shim_check passes
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(vel.tobytes())
        if sigma is not None:
            f.write(sigma.tobytes())
    n = cfg["n"]
    args = [shim_exe, "project", fin, fout] + [str(x) for x in n] + [repr(cfg["dx"][0])] + \
           [str(b) for b in cfg["bclo"]] + [str(b) for b in cfg["bchi"]] + ["1" if sigma is not None else "0"]
    out = subprocess.run(args, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = np.fromfile(fout)
    nv, nn, nc = vel.size, (n[0] + 1) * (n[1] + 1) * (n[2] + 1), 3 * n[0] * n[1] * n[2]
    gvel = raw[:nv].reshape(vel.shape)
    gphi = raw[nv:nv + nn].reshape(n[2] + 1, n[1] + 1, n[0] + 1)
    ggrad = raw[nv + nn:nv + nn + nc].reshape(3, n[2], n[1], n[0])
    iters = int(raw[-1])
    p = oracle_params(cfg["n"], cfg["dx"], cfg["bclo"], cfg["bchi"], tile=(64, 16, 64))
    ovel = vel.copy()
    r = oracle.project(p, ovel, 1, sigma, cfg["const_sigma"], 1e-11, 1e-14)
    assert iters == r["stats"].iters
    assert rel_l2(remove_mean(gphi), remove_mean(r["phi"])) < 1e-9
    assert rel_l2(ggrad, r["gphi"]) < 1e-9
    inner = (slice(None), slice(1, -1), slice(1, -1), slice(1, -1))
    assert rel_l2(gvel[inner], ovel[inner]) < 1e-9


@pytest.mark.gpu
def test_shim_mac_projector_call_sequence(shim_exe):
    """Hydro::MacProjector through the C++ mirror (include/B200MacProjector.H): initProjector(const beta) / setDomainBC /
    project, updateCoeffs(face arrays) / project(mac_phi, ...), warm start"""
    out = subprocess.run([shim_exe, "mac", "32"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shim mac OK" in out.stdout


@pytest.mark.gpu
def test_shim_multibox_apply_equals_single_box(shim_exe):
    """incflo::ApplyNodalProjection through the C++ mirror on a MultiFab of 16^3 boxes (what mfab_of(amrex::MultiFab&) hands
    over with amr.max_grid_size = 16) against the same call on one box: bit-identical"""
    out = subprocess.run([shim_exe, "multibox", "48", "16"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shim multibox OK: 27 boxes" in out.stdout


@pytest.mark.gpu
def test_shim_composite_multibox_apply_equals_single_box(shim_exe):
    """incflo::ApplyNodalProjection with finest_level = 1 through the C++ mirror (IncfloCompositeProjection): both AMR levels as
    MultiFabs of 16^3 boxes against one box per level: bit-identical"""
    out = subprocess.run([shim_exe, "composite_mf", "32", "16"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shim composite_mf OK: 8 + 8 boxes" in out.stdout


@pytest.mark.gpu
def test_shim_two_level_project_matches_oracle(shim_exe, tmp_path, oracle):
    """Hydro::NodalProjector with two-element vectors (finest_level = 1) through the C++ mirror, against the
    composite oracle: periodic x/y, walls z, central box (bouss_bubble-like, BASELINE configs[3] scaled down)"""
    from oracle import composite as oc
    N = 16
    n0, dx = (N, N, N), 1.0 / N
    bclo = bchi = (0, 0, 1)
    flo, fhi = (4, 4, 4), (11, 11, 11)
    nf = (16, 16, 16)
    rng = np.random.default_rng(21)

    def field(n):
        v = rng.standard_normal((3,) + n[::-1])
        for ax in (1, 2, 3):
            v = 0.5 * v + 0.25 * (np.roll(v, 1, ax) + np.roll(v, -1, ax))
        out = np.zeros((3, n[2] + 2, n[1] + 2, n[0] + 2)); out[:, 1:-1, 1:-1, 1:-1] = v
        return out
    vel0, vel1 = field(n0), field(nf)
    fin, fout = str(tmp_path / "cin.bin"), str(tmp_path / "cout.bin")
    with open(fin, "wb") as f:
        f.write(vel0.tobytes()); f.write(vel1.tobytes())
    args = [shim_exe, "composite", fin, fout] + [str(x) for x in n0] + [repr(dx)] + [str(b) for b in bclo] + \
           [str(b) for b in bchi] + [str(x) for x in flo] + [str(x) for x in fhi]
    out = subprocess.run(args, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = np.fromfile(fout)
    sizes = [vel0.size, vel1.size, (N + 1) ** 3, 17 ** 3, 3 * N ** 3, 3 * 16 ** 3, 1]
    assert raw.size == sum(sizes)
    parts = np.split(raw, np.cumsum(sizes)[:-1])
    gv0 = parts[0].reshape(vel0.shape); gv1 = parts[1].reshape(vel1.shape)
    gp0 = parts[2].reshape(N + 1, N + 1, N + 1); gp1 = parts[3].reshape(17, 17, 17)
    gg0 = parts[4].reshape(3, N, N, N); gg1 = parts[5].reshape(3, 16, 16, 16)
    iters = int(parts[6][0])
    cp = oc.CompositeProjector(oracle_params(n0, (dx,) * 3, bclo, bchi, tile=(64, 16, 64)), flo, fhi,
                               smoother_kw=dict(smoother=oracle.SM_BOX, box=(64, 16, 64), box_order=oracle.SM_PLANE4, box_stale_per_call=0))
    ov0, ov1 = vel0.copy(), vel1.copy()
    r = cp.project(ov0, 1, ov1, 1, const_sigma=0.37, rtol=1e-11, atol=1e-14)
    assert r["status"] == 0 and abs(iters - r["iters"]) <= 1
    c = gp1.mean() - r["phi1"].mean()
    p0 = r["phi0"]
    for ax in (1, 2):   # periodic x, y: append the image plane (the mirror returns the full nodal box)
        p0 = np.concatenate([p0, np.take(p0, [0], axis=ax)], axis=ax)
    assert rel_l2(gp1 - c, r["phi1"]) < 1e-9 and rel_l2(gp0 - c, p0) < 1e-9
    assert rel_l2(gg1, r["gphi1"]) < 1e-9 and rel_l2(gg0, r["gphi0"]) < 1e-9
    inner = (slice(None), slice(1, -1), slice(1, -1), slice(1, -1))
    assert rel_l2(gv0[inner], ov0[inner]) < 1e-9 and rel_l2(gv1[inner], ov1[inner]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["eb_channel_cylinder", "eb_cylinder_ebflow"])
def test_shim_eb_project_matches_oracle(shim_exe, tmp_path, name):
    """b200::EBNodalProjector / set_eb_velocity / getLinOp().setEBInflowVelocity (include/B200EBNodalProjector.H) from plain C++"""
    from oracle import eb_oracle as eo
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    n = tuple(int(x) for x in g["n"])
    sigma = float(g["sigma"])
    ebflow = 0.7 if g["eb_vel"].size else 0.0
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        for a in (g["vel"], g["vfrac"], g["intg"], g["bnorm"], g["bintg"]):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    args = [shim_exe, "eb", fin, fout] + [str(x) for x in n] + [repr(float(g["dx"][0]))] + [str(int(b)) for b in g["bclo"]] + \
           [str(int(b)) for b in g["bchi"]] + [repr(sigma), repr(ebflow)]
    out = subprocess.run(args, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = np.fromfile(fout)
    nv, nn, nc = g["vel"].size, (n[0] + 1) * (n[1] + 1) * (n[2] + 1), 3 * n[0] * n[1] * n[2]
    gvel = raw[:nv].reshape(g["vel"].shape)
    gphi = raw[nv:nv + nn].reshape(n[2] + 1, n[1] + 1, n[0] + 1)
    ggrad = raw[nv + nn:nv + nn + nc].reshape(3, n[2], n[1], n[0])
    gebv = raw[nv + nn + nc:nv + nn + nc + nv].reshape(g["vel"].shape)
    p = eo.Params(n, tuple(g["dx"]), g["bclo"], g["bchi"])
    ebv = g["eb_vel"] if g["eb_vel"].size else None
    ref = eo.project(p, g["vel"], sigma, g["vfrac"], g["intg"], 1e-11, 1e-14, ebv, g["bnorm"], g["bintg"])
    assert int(raw[-1]) == ref["info"]["iters"]
    assert rel_l2(gvel[:, 1:-1, 1:-1, 1:-1], ref["vel"]) < 1e-9 and rel_l2(ggrad, ref["gphi"]) < 1e-9
    assert rel_l2(gphi[: ref["phi"].shape[0]], ref["phi"]) < 1e-9      # periodic z: the duplicate top plane is extra
    if ebflow:
        assert np.array_equal(gebv[:, 1:-1, 1:-1, 1:-1], g["eb_vel"])
