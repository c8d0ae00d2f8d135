"""N2: g2o ingest / egress of the harness (CPU only)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
g2o = importlib.import_module("kimera-rpgo_b200.g2o")
synth = importlib.import_module("kimera-rpgo_b200.synth")


def test_write_then_load_3d_round_trip(tmp_path):
    gph = synth.config2(seed=2, P=200, n=12)
    edges = [(f[1], f[2], f[3], f[4]) for f in gph["odom"] + gph["lcs"]]
    path = str(tmp_path / "graph.g2o")
    g2o.write_g2o(path, gph["values"], edges, d=3)
    lines = open(path).read().splitlines()
    assert lines[0].startswith("VERTEX_SE3:QUAT ") and len(lines[0].split()) == 9
    e0 = [l for l in lines if l.startswith("EDGE_SE3:QUAT ")][0].split()
    assert len(e0) == 3 + 7 + 21
    values, back = g2o.load3d(path)
    assert [v[0] for v in values] == [v[0] for v in gph["values"]]
    for (k1, k2, p, c), (b1, b2, bp, bc) in zip(edges, back):
        assert (k1, k2) == (b1, b2)
        np.testing.assert_allclose(bp, p, atol=1e-12)
        np.testing.assert_allclose(bc, c, rtol=1e-9, atol=1e-15)
    # g2o order is translation-first: the first information entry of an odometry edge is 1/sigma_t^2 = 1000
    assert abs(float(e0[10]) - 1000.0) < 1e-6


def test_write_then_load_2d_round_trip(tmp_path):
    gph = synth.config3(seed=2, P=200, n=8)
    edges = [(f[1], f[2], f[3], f[4]) for f in gph["odom"] + gph["lcs"]]
    path = str(tmp_path / "graph2d.g2o")
    g2o.write_g2o(path, gph["values"], edges, d=2)
    values, back = g2o.load2d(path)
    assert len(values) == 200 and len(back) == len(edges)
    for (k1, k2, p, c), (b1, b2, bp, bc) in zip(edges, back):
        np.testing.assert_allclose(bp, p, atol=1e-12)
        np.testing.assert_allclose(bc, c, rtol=1e-9)
