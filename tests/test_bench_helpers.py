"""bench.py host helpers that need no GPU: the whole-solve algorithmic byte count and the
peak lookup."""
import os
import sys

from conftest import ROOT

sys.path.insert(0, ROOT)


def test_solve_roofline_matches_survey_table():
    import bench
    nodes, cells = 257 * 256 * 256, 256 ** 3
    b = bench.solve_roofline(nodes, cells, 7, True)
    # 725 B/node/V-cycle (634 x 8/7) x 7 + 32 B/node + 184 B/cell  ("bytes_solve")
    assert abs(b - (nodes * (634 * 8 / 7 * 7 + 32) + cells * 184)) < 1e-6 * b
    assert 88e9 < b < 90e9
    assert bench.solve_roofline(nodes, cells, 7, False) < b      # constant sigma moves fewer bytes


def test_peak_lookup_has_a_source():
    import bench
    peak, src = bench.peaks()
    assert peak > 1000 and ("measured" in src or "fallback" in src)
