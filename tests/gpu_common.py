"""Helpers for the GPU parity tests: load the product package (hyphenated directory name)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pkg = importlib.import_module("kimera-rpgo_b200")
synth = importlib.import_module("kimera-rpgo_b200.synth")
PcmGpu = pkg.PcmGpu


class ThreadedExchange:
    """In-process stand-in for the collective behind rpgo_set_exchange: `world` threads (one handle each, same GPU)
    meet at a barrier.  Lets the single-GPU suite run the rank-partitioned clique searches end to end."""

    def __init__(self, world):
        import threading
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.calls = 0

    def callback(self, rank):
        import ctypes as C

        import numpy as np
        capi = pkg._capi

        def fn(_user, op, buf, count, root):
            ctype = C.c_int32 if op == capi.XCHG_BCAST_I32 else C.c_int64
            arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(ctype)), shape=(count,))
            self.slots[rank] = arr.copy()
            self.barrier.wait()
            if op == capi.XCHG_MIN_I64:
                out = np.min(np.stack(self.slots), axis=0)
            elif op == capi.XCHG_MAX_I64:
                out = np.max(np.stack(self.slots), axis=0)
            else:
                out = self.slots[root]
            arr[:] = out
            if rank == 0:
                self.calls += 1
            self.barrier.wait()
            return 0

        return capi.EXCHANGE_FN(fn)


def run_sharded(world, make_handle, work):
    """make_handle(rank) -> PcmGpu with cfg.rank/world set; work(handle) -> result.  Returns the per-rank results."""
    import threading
    ex = ThreadedExchange(world)
    out, err = [None] * world, [None] * world

    def body(r):
        try:
            h = make_handle(r)
            h.set_exchange(ex.callback(r))
            out[r] = work(h)
            h.close()
        except BaseException as e:  # noqa: BLE001
            err[r] = e
            ex.barrier.abort()

    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    for e in err:
        if e is not None:
            raise e
    return out, ex.calls
