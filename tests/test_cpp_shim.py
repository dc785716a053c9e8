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
