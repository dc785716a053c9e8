"""IncfloVelFill (src/prob/prob_bc.H:8-351): the host restatement against hand-computed values (CPU), and the
device kernel behind b200np_set_inflow_profile against the oracle fed with the host restatement (GPU)."""
import numpy as np
import pytest

from helpers import TILE, oracle_params, rel_l2

INF, NEU, DIR, PER = 3, 1, 2, 0


def test_host_restatement_matches_reference_formulas():
    from incflo_b200 import prob_bc
    n = (8, 4, 6)
    bcv = np.zeros((6, 3)); bcv[0] = (1.0, 0.25, -0.5)          # x-lo: (normal, tangential y, tangential z)
    # channel_cylinder-like deck: probtype 31, parabolic profile 6 y (1 - y) on x-lo (prob_bc.H:62-66)
    a = prob_bc.incflo_vel_fill(n, 2, (INF, NEU, PER), (DIR, NEU, PER), 31, bcv)
    for j in range(4):
        y = (j + 0.5) / 4
        assert np.allclose(a[0, 2:-2, 2 + j, 1], 6 * y * (1 - y))
    assert np.allclose(a[1, 2:-2, 2:-2, 1], 0.25) and np.allclose(a[2, 2:-2, 2:-2, 1], -0.5)
    assert np.count_nonzero(a[:, :, :, 2:]) == 0 and np.count_nonzero(a[:, :, :, 0]) == 0   # only the first ghost layer of x-lo
    # probtype 42: normal velocity = time on both x faces (:57-60, :129-132)
    bcv[3] = (9.0, 0.0, 0.0)
    a = prob_bc.incflo_vel_fill(n, 1, (INF, NEU, NEU), (INF, NEU, NEU), 42, bcv, time=0.375)
    assert np.allclose(a[0, 1:-1, 1:-1, 0], 0.375) and np.allclose(a[0, 1:-1, 1:-1, -1], 0.375)
    # probtype 16 on y-hi (:241-246) and the order of the face blocks at a corner ghost cell (x-lo then y-hi)
    bcv = np.zeros((6, 3)); bcv[0] = (2.0, 3.0, 4.0); bcv[4] = (5.0, 6.0, 7.0)
    a = prob_bc.incflo_vel_fill(n, 1, (INF, NEU, NEU), (NEU, INF, NEU), 16, bcv)
    x = (np.arange(8) + 0.5) / 8
    assert np.allclose(a[1, 3, -1, 1:-1], 16 * (x ** 4 - 2 * x ** 3 + x ** 2))
    assert np.allclose(a[0, 3, -1, 1:-1], 5.0) and np.allclose(a[2, 3, -1, 1:-1], 7.0)
    xg = (-1 + 0.5) / 8                                      # corner (i = -1, j = ny): the y-hi block wrote last
    assert np.isclose(a[1, 3, -1, 0], 16 * (xg ** 4 - 2 * xg ** 3 + xg ** 2)) and np.isclose(a[0, 3, -1, 0], 5.0)
    assert np.isclose(a[0, 3, 2, 0], 2.0) and np.isclose(a[1, 3, 2, 0], 3.0)
