#!/usr/bin/env python
"""Regenerate tests/golden/g2o_fixtures.npz from the reference's own g2o fixtures.

/root/reference does not exist on the GPU box, so the four pose graphs the reference's
integration tests load (tests/data/{robot_a,robot_b,ordered,unordered}.g2o, used by
tests/testLoadGraph.cpp et al.) are parsed HERE, in the build container, with the harness's g2o
reader and stored as arrays (keys, poses, covariances in GTSAM tangent order).
Run:  python tools/make_golden_from_reference.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
g2o = importlib.import_module("kimera-rpgo_b200.g2o")

REF = "/root/reference/tests/data"
out = {}
for name in ["robot_a", "robot_b", "ordered", "unordered"]:
    values, edges = g2o.load3d(os.path.join(REF, name + ".g2o"))
    out[name + "_vkeys"] = np.array([v[0] for v in values], dtype=np.uint64)
    out[name + "_vposes"] = np.array([v[1] for v in values], dtype=np.float64)
    out[name + "_k1"] = np.array([e[0] for e in edges], dtype=np.uint64)
    out[name + "_k2"] = np.array([e[1] for e in edges], dtype=np.uint64)
    out[name + "_eposes"] = np.array([e[2] for e in edges], dtype=np.float64)
    out[name + "_ecovs"] = np.array([e[3] for e in edges], dtype=np.float64)
    print(name, len(values), "vertices", len(edges), "edges")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "g2o_fixtures.npz"), **out)
