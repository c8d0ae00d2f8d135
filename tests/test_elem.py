"""include/rpgo_elem.h: deterministic elementary functions vs mpmath (<= 3 ulp on the domains the path uses)."""
import ctypes as C
import math
import os
import random
import subprocess

import pytest

mp = pytest.importorskip("mpmath")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "_elem_shim.so")
SRC = """
#include "rpgo_elem.h"
double t_sin(double x){return rpgo_sin(x);} double t_cos(double x){return rpgo_cos(x);}
double t_tan(double x){return rpgo_tan(x);} double t_acos(double x){return rpgo_acos(x);}
double t_atan2(double y,double x){return rpgo_atan2(y,x);}
"""


@pytest.fixture(scope="module")
def lib():
    hdr = os.path.join(ROOT, "include", "rpgo_elem.h")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(hdr):
        c = os.path.join(ROOT, "tests", "_elem_shim.c")
        with open(c, "w") as f:
            f.write(SRC)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                               c, "-o", SO, "-lm"])
        os.remove(c)
    L = C.CDLL(SO)
    for n in ["t_sin", "t_cos", "t_tan", "t_acos"]:
        getattr(L, n).restype = C.c_double
        getattr(L, n).argtypes = [C.c_double]
    L.t_atan2.restype = C.c_double
    L.t_atan2.argtypes = [C.c_double, C.c_double]
    return L


def ulps(got, ref):
    r = float(ref)
    if r == 0:
        return abs(got) / 5e-324
    return float(abs(mp.mpf(got) - ref) / mp.mpf(math.ulp(r)))


def test_elementary_functions_within_3_ulp(lib):
    mp.mp.dps = 40
    random.seed(3)
    worst = {}
    for name, f, ref, dom in [("sin", lib.t_sin, mp.sin, (-3.3, 3.3)), ("cos", lib.t_cos, mp.cos, (-3.3, 3.3)),
                              ("tan", lib.t_tan, mp.tan, (0.0, 1.58)), ("acos", lib.t_acos, mp.acos, (-1.0, 1.0))]:
        m = 0.0
        for i in range(20000):
            x = random.uniform(*dom)
            if name == "acos" and i % 4 == 0:
                x = 1 - 10 ** random.uniform(-12, 0)
            if name == "acos" and i % 4 == 1:
                x = -1 + 10 ** random.uniform(-12, 0)
            m = max(m, ulps(f(x), ref(mp.mpf(x))))
        worst[name] = m
    m = 0.0
    for i in range(20000):
        y = random.uniform(-1, 1) * 10 ** random.uniform(-8, 2)
        x = random.uniform(-1, 1) * 10 ** random.uniform(-8, 2)
        m = max(m, ulps(lib.t_atan2(y, x), mp.atan2(mp.mpf(y), mp.mpf(x))))
    worst["atan2"] = m
    assert all(v <= 3.0 for v in worst.values()), worst


def test_special_values(lib):
    assert lib.t_acos(1.0) == 0.0 and lib.t_acos(-1.0) == math.pi and lib.t_acos(0.0) == math.pi / 2
    assert math.isnan(lib.t_acos(1.0000001))
    assert lib.t_atan2(0.0, -1.0) == math.pi and lib.t_atan2(-0.0, -1.0) == -math.pi
    assert lib.t_atan2(1.0, 0.0) == math.pi / 2 and lib.t_atan2(0.0, 1.0) == 0.0
    assert lib.t_sin(0.0) == 0.0 and lib.t_cos(0.0) == 1.0
