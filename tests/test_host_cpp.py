"""The C++ host API (kimera-rpgo_b200/host/rpgo_host.hpp): compiles on CPU, runs on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_test_host_api")


def build():
    src = os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp")
    hdr = os.path.join(ROOT, "kimera-rpgo_b200", "host", "rpgo_host.hpp")
    lib = os.path.join(ROOT, "kimera-rpgo_b200", "librpgo_b200.so")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(lib)):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "kimera-rpgo_b200", "host"), src,
                               "-L", os.path.join(ROOT, "kimera-rpgo_b200"), "-lrpgo_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "kimera-rpgo_b200"), "-Wl,-rpath,$ORIGIN/../../kimera-rpgo_b200",
                               "-o", EXE])
    return EXE


def test_host_api_compiles_and_fails_loudly_without_gpu():
    exe = build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_host_api_reference_scenarios_on_gpu():
    exe = build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "There were no test failures" in r.stdout
