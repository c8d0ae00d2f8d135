"""Host-side pieces that need no GPU: the flat key table of the C-ABI library (compiled for the host) and the lazy id
lists of the Python mirror."""
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_keymap_against_unordered_map(tmp_path):
    exe = str(tmp_path / "test_keymap")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "kimera-rpgo_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "test_keymap.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "KEYMAP PASS" in r.stdout, r.stdout + r.stderr


def test_idlist_behaves_like_a_list():
    IdList = importlib.import_module("kimera-rpgo_b200.pcm").IdList
    ref = []
    x = IdList()
    assert len(x) == 0 and not x and list(x) == []
    x.extend(range(3, 8)); ref.extend(range(3, 8))
    x.extend(np.arange(100, 104, dtype=np.int32)); ref.extend([100, 101, 102, 103])
    assert len(x) == len(ref) == 9 and bool(x)
    # integer indexing resolves inside the deferred chunks (no materialisation) and returns Python ints
    assert [x[i] for i in range(9)] == ref and x[-1] == 103 and isinstance(x[6], int)
    x.append(7); ref.append(7)
    x.extend([1, 2]); ref.extend([1, 2])
    x.extend(np.array([], dtype=np.int64))
    assert list(x) == ref and x == ref and len(x) == len(ref)
    assert x.pop() == ref.pop() and x.index(101) == ref.index(101)
    x.extend(np.arange(5)); ref.extend(range(5))
    assert x.pop() == ref.pop() and len(x) == len(ref) and x[len(ref) - 1] == ref[-1]
    assert all(isinstance(v, int) for v in x) and list(x) == ref
    y = IdList([9, 8])
    y.extend(np.arange(3, dtype=np.int32))
    y.extend(range(20, 22))
    assert np.array(y, dtype=np.int64).tolist() == [9, 8, 0, 1, 2, 20, 21] and np.asarray(y).dtype == np.int64
    assert len(y) == 7 and y[2] == 0  # conversion does not consume the chunks
    try:
        IdList(range(2))[5]
        assert False
    except IndexError:
        pass
