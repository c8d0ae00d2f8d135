"""world_size-2 gloo test (CPU) of the one exchange step of the multi-GPU path: the all-gather of adjacency
row chunks (kimera-rpgo_b200/parallel.py::exchange_row_chunks)."""
import importlib
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, chunk_rows, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = importlib.import_module("kimera-rpgo_b200.parallel")
    rng = np.random.default_rng(5)
    padded = 2 * world * chunk_rows
    words = (n + 63) // 64
    full = torch.from_numpy(rng.integers(-2**62, 2**62, size=(padded, words), dtype=np.int64))
    mine = torch.zeros_like(full)
    for c in par.owned_chunks(rank, world):
        mine[c * chunk_rows:(c + 1) * chunk_rows] = full[c * chunk_rows:(c + 1) * chunk_rows]
    out = par.exchange_row_chunks(mine, rank, world, chunk_rows)
    q.put((rank, bool(torch.equal(out, full))))
    dist.destroy_process_group()


def test_exchange_row_chunks_world2():
    world, n, chunk_rows = 2, 1000, 256
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, chunk_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_owned_chunks_balance_the_triangle():
    """chunks r and 2W-1-r together hold (almost) the same number of upper-triangle pairs for every rank."""
    par = importlib.import_module("kimera-rpgo_b200.parallel")
    n, world = 50000, 8
    chunk = -(-n // (2 * world))
    chunk = (chunk + 31) // 32 * 32
    work = []
    for r in range(world):
        w = 0
        for c in par.owned_chunks(r, world):
            lo, hi = c * chunk, min((c + 1) * chunk, n)
            rows = np.arange(lo, max(hi, lo))
            w += int((n - 1 - rows).sum())
        work.append(w)
    assert max(work) / (sum(work) / world) < 1.02


def _xchg_worker(rank, world, port, q):
    import ctypes as C
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = importlib.import_module("kimera-rpgo_b200.parallel")
    capi = importlib.import_module("kimera-rpgo_b200._capi")
    cb = par.make_exchange(None)  # CPU tensors over the process group
    key = (C.c_int64 * 1)((7 + rank) << 32 | (100 - rank))
    rc1 = cb(None, capi.XCHG_MIN_I64, C.cast(key, C.c_void_p), 1, 0)
    lo = key[0]
    none = (C.c_int64 * 1)(2**63 - 1 if rank == 0 else 5)   # "no improver" on rank 0 must lose against any key
    rc2 = cb(None, capi.XCHG_MIN_I64, C.cast(none, C.c_void_p), 1, 0)
    inc = (C.c_int64 * 1)((3 + rank) << 32 | 9)
    rc3 = cb(None, capi.XCHG_MAX_I64, C.cast(inc, C.c_void_p), 1, 0)
    ids = (C.c_int32 * 4)(*([11, 12, 13, 14] if rank == 1 else [0, 0, 0, 0]))
    rc4 = cb(None, capi.XCHG_BCAST_I32, C.cast(ids, C.c_void_p), 4, 1)
    q.put((rank, rc1 | rc2 | rc3 | rc4, lo, none[0], inc[0], list(ids)))
    dist.destroy_process_group()


def test_clique_incumbent_exchange_world2():
    """the collective behind rpgo_set_exchange (all-reduce MIN/MAX of the packed incumbent, broadcast of the ids)"""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_xchg_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for rank, rc, lo, none, inc, ids in res:
        assert rc == 0
        assert lo == (7 << 32 | 100)          # lowest candidate index wins the round
        assert none == 5
        assert inc == (4 << 32 | 9)           # larger clique wins
        assert ids == [11, 12, 13, 14]
