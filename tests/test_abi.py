"""CPU-only: the C-ABI library loads and exports every symbol include/rpgo_b200.h declares; using it without
a GPU fails loudly (no CPU fallback); the product never imports the oracle."""
import ctypes as C
import os
import re

import pytest

from gpu_common import pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rpgo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rpgo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = pkg._capi.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    bound = {s[0] for s in pkg._capi.SYMBOLS}
    assert set(names) == bound, set(names) ^ bound


def test_cfg_struct_matches_header_defaults():
    lib = pkg._capi.load()
    cfg = pkg._capi.RpgoCfg()
    assert lib.rpgo_default_cfg(C.byref(cfg)) == 0
    # PcmParams defaults, reference SolverParams.h:35-42
    assert (cfg.odom_threshold, cfg.lc_threshold) == (10.0, 5.0)
    assert (cfg.odom_trans_threshold, cfg.odom_rot_threshold) == (0.05, 0.005)
    assert (cfg.dist_trans_threshold, cfg.dist_rot_threshold) == (0.01, 0.001)
    assert cfg.world == 1 and cfg.band == 1e-9


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.RpgoError):
        pkg.PcmGpu(3, 0)


def test_product_never_references_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "kimera-rpgo_b200")):
        if os.path.basename(base) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(base, f)).read()
                assert "liboracle" not in txt and "import orc" not in txt and "oracle/" not in txt, os.path.join(base, f)
    assert "oracle" not in open(os.path.join(ROOT, "include", "rpgo_b200.h")).read()
