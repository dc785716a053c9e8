"""Shared helpers of the parity tests: BC cases, random fields, oracle set-up that mirrors the
GPU smoother (Gauss-Seidel inside 64x16xTZ tiles, previous-sweep values outside)."""
import numpy as np

from oracle import pyoracle as po

TILE = (64, 16, 16)

# (name, n_cell, dx, bclo, bchi)
BC_CASES = [
    ("periodic", (32, 32, 32), (1 / 32,) * 3, (0, 0, 0), (0, 0, 0)),
    ("rt_walls_z", (32, 16, 24), (1 / 32,) * 3, (0, 0, 1), (0, 0, 1)),
    ("all_neumann", (16, 24, 16), (1 / 16,) * 3, (1, 1, 1), (1, 1, 1)),
    ("channel_inflow_outflow", (48, 16, 16), (1 / 48, 1 / 48, 1 / 48), (3, 1, 0), (2, 1, 0)),
    ("anisotropic_dirichlet_z", (16, 16, 32), (0.1, 0.07, 0.05), (0, 1, 2), (0, 1, 2)),
    ("wide_two_tiles", (80, 40, 8), (1 / 80,) * 3, (0, 1, 0), (0, 1, 0)),
]


def oracle_params(n, dx, bclo, bchi, tile=TILE, **kw):
    """oracle configured like the GPU smoother"""
    base = dict(smoother=po.SM_BOX, box=tile, box_order=po.SM_PLANE4, box_stale_per_call=0)
    base.update(kw)
    return po.make_params(n, dx, bclo, bchi, **base)


def random_sigma(n, rng, contrast=4.0):
    return np.ascontiguousarray(rng.uniform(1.0, contrast, size=(n[2], n[1], n[0])))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm((a - b).ravel())
    s = np.linalg.norm(b.ravel())
    return d / s if s > 0 else d


def remove_mean(a):
    return a - a.mean()
