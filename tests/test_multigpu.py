"""Multi-GPU runs exactly as the driver launches them (one process per GPU under torch.distributed.run), collected by
pytest so that they run on any box with >= 2 GPUs:
  * tests/mgpu_check.py   sharded == single GPU (bitset, degrees, flagged pairs, inliers), the three clique modes, and
                          1e6 sampled pairs of the 50k-closure all-gathered matrix against the CPU oracle;
  * bench.py              must print its JSON line AND exit 0 (round 1 crashed at teardown for every N > 1);
  * tests/cpp/test_comm_ranks.cpp   the same data plane driven from C++ only (fork + pipes, no Python).
"""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(n, script, *args, env=None, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), script] + list(args)
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=e)


needs2 = pytest.mark.skipif(_gpus() < 2, reason="needs >= 2 GPUs")


@pytest.mark.gpu
@needs2
def test_mgpu_check_under_torchrun():
    n = 2
    r = _torchrun(n, os.path.join(ROOT, "tests", "mgpu_check.py"), env={"MGPU_BIG_SAMPLE": "1000000"})
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "MGPU_CHECK PASS" in r.stdout


@pytest.mark.gpu
@needs2
@pytest.mark.parametrize("n", [2, 4, 8])
def test_bench_exits_zero_under_torchrun(n):
    """the driver's own invocation (incl. NCCL_DEBUG=INFO to a file), smaller workload"""
    if _gpus() < n:
        pytest.skip("needs %d GPUs" % n)
    env = {"NCCL_DEBUG": "INFO", "NCCL_DEBUG_FILE": "/tmp/nccl.%p.log"}
    r = _torchrun(n, os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--steps", "3", "--warmup", "3", "--closures", "12000",
                  "--poses", "12000", "--extras", "config4", env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["n_gpus"] == n and out["value"] > 0 and out["gpu_launches"] > 0
    assert out["e2e"]["h2d_bytes_per_step"] > 0
    assert out["extra"]["config4"]["groups"] == 36


def _build_cpp():
    exe = os.path.join(ROOT, "tests", "cpp", "_test_comm_ranks")
    src = os.path.join(ROOT, "tests", "cpp", "test_comm_ranks.cpp")
    lib = os.path.join(ROOT, "kimera-rpgo_b200", "librpgo_b200.so")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src,
                               "-L", os.path.join(ROOT, "kimera-rpgo_b200"), "-lrpgo_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "kimera-rpgo_b200"), "-Wl,-rpath,$ORIGIN/../../kimera-rpgo_b200",
                               "-o", exe])
    return exe


def test_cpp_rank_test_compiles_and_skips_without_gpus():
    exe = _build_cpp()
    if _gpus() >= 2:
        pytest.skip("GPUs present: covered by the gpu test")
    r = subprocess.run([exe, "2", "500"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 77 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
@needs2
def test_cpp_two_ranks_without_python():
    exe = _build_cpp()
    r = subprocess.run([exe, "2", "3000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "COMM_RANKS PASS" in r.stdout
