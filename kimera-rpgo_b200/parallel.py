"""Multi-GPU plumbing around the C ABI.

The data plane itself is C++/NCCL inside the library (csrc/comm.cu: rpgo_comm_init, the all-gather of adjacency row
chunks inside rpgo_lc_append, the incumbent all-reduce of the clique searches).  What lives here is
  * `comm_id_via_torch` / `attach_comm`: hand rank 0's NCCL unique id to the other ranks over an existing
    torch.distributed process group (one process per GPU) and build the handle's communicator;
  * `exchange_row_chunks`: the same exchange protocol written with torch.distributed collectives.  It mirrors
    comm.cu::comm_allgather_row_chunks step by step (in-place all-gather of the low chunks, one broadcast per rank for
    the mirrored high chunks) and runs on CPU tensors with the gloo backend, which is how the chunk layout is tested
    without GPUs;
  * `make_exchange`: a host collective for rpgo_set_exchange (bring-your-own transport; gloo in the tests).
Every rank holds the replicated trajectory and closure tables and computes the upper-triangle pair checks of its row
chunks {rank, 2*world-1-rank} (pairing a long and a short chunk balances the triangle)."""
import torch
import torch.distributed as dist


def comm_id_via_torch(unique_id_fn, group=None):
    """rank 0 calls unique_id_fn() (PcmGpu.comm_unique_id); returns the same 128 bytes on every rank."""
    box = [unique_id_fn() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0 if group is None else dist.get_global_rank(group, 0), group=group)
    return box[0]


def attach_comm(pcm, group=None):
    """Collective over the process group: give `pcm` (created with rank/world of this group) its NCCL communicator."""
    pcm.comm_init(comm_id_via_torch(type(pcm).comm_unique_id, group))
    return pcm


def owned_chunks(rank, world):
    return (rank, 2 * world - 1 - rank)


def exchange_row_chunks(t, rank, world, chunk_rows, group=None):
    """t: (2*world*chunk_rows, row_words) integer tensor whose rows in this rank's two chunks are valid.
    After the call every rank holds all rows.  Same sequence as comm.cu: one in-place all-gather of chunks 0..world-1
    (rank r's chunk sits at slot r), then for every rank q a broadcast of chunk 2*world-1-q from q."""
    if world == 1:
        return t
    lo = [t[q * chunk_rows:(q + 1) * chunk_rows] for q in range(world)]
    mine_lo = lo[rank].clone()
    dist.all_gather(lo, mine_lo, group=group)
    for q in range(world):
        c = 2 * world - 1 - q
        src = q if group is None else dist.get_global_rank(group, q)
        dist.broadcast(t[c * chunk_rows:(c + 1) * chunk_rows], src=src, group=group)
    return t


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
