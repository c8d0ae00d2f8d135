"""Multi-GPU plumbing for the one exchange step the path has: the all-gather of adjacency row chunks.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Every rank holds the replicated
trajectory and closure tables and computes the upper-triangle pair checks of its row chunks
{rank, 2*world-1-rank} (pairing a long and a short chunk balances the triangle).  The chunks are then
all-gathered in place and every rank rebuilds the lower triangle and the degrees locally
(rpgo_group_finalize).  The same function runs on CPU tensors with the gloo backend for the tests."""
import torch
import torch.distributed as dist


class _DevMem:
    """Zero-copy view of library-owned device memory for torch (via __cuda_array_interface__)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def tensor_from_device_ptr(ptr, nbytes, device):
    return torch.as_tensor(_DevMem(ptr, nbytes), device=device)


def owned_chunks(rank, world):
    return (rank, 2 * world - 1 - rank)


def exchange_row_chunks(t, rank, world, chunk_rows, group=None):
    """t: (2*world*chunk_rows, row_words) integer tensor whose rows in this rank's two chunks are valid.
    After the call every rank holds all rows.  Two all-gathers of equal-size chunks."""
    if world == 1:
        return t
    lo = [t[q * chunk_rows:(q + 1) * chunk_rows] for q in range(world)]
    hi = [t[(2 * world - 1 - q) * chunk_rows:(2 * world - q) * chunk_rows] for q in range(world)]
    mine_lo = lo[rank].clone()
    mine_hi = hi[rank].clone()
    dist.all_gather(lo, mine_lo, group=group)
    dist.all_gather(hi, mine_hi, group=group)
    return t


def allgather_adjacency(pcm, g, device, group=None):
    """All-gather group g's adjacency rows across ranks on the handle's stream, then finalize."""
    world = pcm.cfg.world
    if world > 1:
        ptr, sw64, n = pcm.adj_bits_device(g)
        chunk_rows, padded = pcm.group_chunking(g)
        t = tensor_from_device_ptr(ptr, padded * sw64 * 8, device).view(torch.int64).view(padded, sw64)
        st = torch.cuda.ExternalStream(pcm.stream_ptr(), device=device)
        with torch.cuda.stream(st):
            exchange_row_chunks(t, pcm.cfg.rank, world, chunk_rows, group=group)
    pcm.finalize(g)


def make_exchange(device=None, group=None):
    """Host collective for the sharded clique searches (rpgo_set_exchange): all-reduce MIN/MAX of int64 and
    broadcast of int32 over torch.distributed — NCCL (device tensors, NVLink) when `device` is a CUDA device,
    the process group's CPU backend otherwise (gloo in the tests).  Returns the ctypes callback; the caller must
    keep a reference to it for as long as it is registered."""
    import ctypes as C

    import numpy as np

    from . import _capi

    state = {}

    def staging(dtype, count):
        key = (dtype, count)
        if key not in state:
            state[key] = (torch.empty(count, dtype=dtype, pin_memory=True), torch.empty(count, dtype=dtype, device=device))
        return state[key]

    def fn(_user, op, buf, count, root):
        try:
            if op == _capi.XCHG_BCAST_I32:
                arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_int32)), shape=(count,))
            else:
                arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_int64)), shape=(count,))
            t = torch.from_numpy(arr)
            if device is not None:
                pin, td = staging(t.dtype, count)
                pin.copy_(t)
                td.copy_(pin, non_blocking=True)
            else:
                td = t
            if op == _capi.XCHG_MIN_I64:
                dist.all_reduce(td, op=dist.ReduceOp.MIN, group=group)
            elif op == _capi.XCHG_MAX_I64:
                dist.all_reduce(td, op=dist.ReduceOp.MAX, group=group)
            elif op == _capi.XCHG_BCAST_I32:
                dist.broadcast(td, src=root if group is None else dist.get_global_rank(group, root), group=group)
            else:
                return 2
            if device is not None:
                pin.copy_(td, non_blocking=True)
                torch.cuda.current_stream(device).synchronize()
                t.copy_(pin)
            return 0
        except Exception as e:  # never unwind through the C frame
            import sys
            print("rpgo exchange callback failed: %r" % (e,), file=sys.stderr)
            return 1

    return _capi.EXCHANGE_FN(fn)
