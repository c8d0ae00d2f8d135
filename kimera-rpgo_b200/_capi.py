"""ctypes binding of include/rpgo_b200.h.  Fails loudly when the CUDA library is missing or no GPU
is usable: there is no CPU fallback in the product."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librpgo_b200.so")

c_dp = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_u64p = C.POINTER(C.c_uint64)


class RpgoCfg(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("mode", C.c_int32),
        ("odom_threshold", C.c_double), ("lc_threshold", C.c_double),
        ("odom_trans_threshold", C.c_double), ("odom_rot_threshold", C.c_double),
        ("dist_trans_threshold", C.c_double), ("dist_rot_threshold", C.c_double),
        ("incremental", C.c_int32), ("device", C.c_int32), ("traj_mode", C.c_int32), ("kernel", C.c_int32),
        ("rank", C.c_int32), ("world", C.c_int32), ("band", C.c_double), ("scan_chunk", C.c_int32),
        ("reserved", C.c_int32),
    ]


# every symbol include/rpgo_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("rpgo_default_cfg", C.c_int, [C.POINTER(RpgoCfg)]),
    ("rpgo_create", C.c_int, [C.POINTER(RpgoCfg), C.POINTER(C.c_void_p)]),
    ("rpgo_destroy", None, [C.c_void_p]),
    ("rpgo_reset", C.c_int, [C.c_void_p]),
    ("rpgo_last_error", C.c_char_p, [C.c_void_p]),
    ("rpgo_sync", C.c_int, [C.c_void_p]),
    ("rpgo_stream", C.c_void_p, [C.c_void_p]),
    ("rpgo_odom_append", C.c_int, [C.c_void_p, C.c_int64, c_u64p, c_u64p, c_dp, c_dp, c_dp]),
    ("rpgo_traj_get", C.c_int, [C.c_void_p, C.c_uint64, c_dp, c_dp, c_i32p, c_i32p]),
    ("rpgo_traj_size", C.c_int64, [C.c_void_p]),
    ("rpgo_lc_append", C.c_int, [C.c_void_p, C.c_int64, c_u64p, c_u64p, c_dp, c_dp, c_u8p, c_i32p, c_i32p, c_dp]),
    ("rpgo_landmark_append", C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, c_u64p, c_dp, c_dp, C.c_int32, c_i32p]),
    ("rpgo_num_groups", C.c_int32, [C.c_void_p]),
    ("rpgo_group_info", C.c_int, [C.c_void_p, C.c_int32, c_u8p, c_u8p, c_i64p]),
    ("rpgo_find_group", C.c_int32, [C.c_void_p, C.c_uint8, C.c_uint8]),
    ("rpgo_lc_remove_last", C.c_int, [C.c_void_p, C.c_int32, c_u64p, c_u64p]),
    ("rpgo_find_inliers", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, c_i32p, c_i64p, c_i32p]),
    ("rpgo_find_inliers_batch", C.c_int, [C.c_void_p, C.c_int32, c_i32p, C.c_int32, c_i64p, c_i64p, c_i32p, c_i64p, c_i64p]),
    ("rpgo_set_exchange", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("rpgo_comm_unique_id", C.c_int, [C.c_void_p]),
    ("rpgo_comm_init", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    ("rpgo_comm_destroy", C.c_int, [C.c_void_p]),
    ("rpgo_group_allgather", C.c_int, [C.c_void_p, C.c_int32]),
    ("rpgo_frame_align_measurements", C.c_int, [C.c_void_p, C.c_int32, C.c_uint8, C.c_int64, c_i32p, c_dp]),
    ("rpgo_robot_odom_values", C.c_int, [C.c_void_p, C.c_uint8, c_dp, C.c_int64, c_u64p, c_dp, c_i64p]),
    ("rpgo_adj_bits", C.c_int, [C.c_void_p, C.c_int32, c_u64p, C.c_int64]),
    ("rpgo_adj_bits_device", C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), c_i64p, c_i64p]),
    ("rpgo_degrees", C.c_int, [C.c_void_p, C.c_int32, c_i32p]),
    ("rpgo_near_threshold", C.c_int, [C.c_void_p, C.c_int32, c_i32p, C.c_int64, c_i64p]),
    ("rpgo_pair_distances", C.c_int, [C.c_void_p, C.c_int32, c_dp]),
    ("rpgo_group_recompute", C.c_int, [C.c_void_p, C.c_int32, C.c_int64]),
    ("rpgo_group_pairwise", C.c_int, [C.c_void_p, C.c_int32, C.c_int64]),
    ("rpgo_group_finalize", C.c_int, [C.c_void_p, C.c_int32]),
    ("rpgo_clique_stats", C.c_int, [C.c_void_p, c_i64p, c_i64p, c_i32p]),
    ("rpgo_debug_pass", C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    ("rpgo_group_chunking", C.c_int, [C.c_void_p, C.c_int32, c_i64p, c_i64p]),
    ("rpgo_launch_count", C.c_int64, [C.c_void_p]),
    ("rpgo_fp64_peak", C.c_int, [C.c_int32, c_dp]),
    ("rpgo_debug_load_group", C.c_int, [C.c_void_p, C.c_uint8, C.c_uint8, C.c_int64, c_u64p, C.c_int64, c_i32p]),
    ("rpgo_debug_check_fastmath", C.c_int, [C.c_int64, C.c_uint64, c_u64p, c_u64p]),
    ("rpgo_version", C.c_char_p, []),
]

EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32)
XCHG_MIN_I64, XCHG_MAX_I64, XCHG_BCAST_I32 = 0, 1, 2
COMM_ID_BYTES = 128

_lib = None


def load():
    """dlopen the in-tree CUDA library.  Raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        # measurement builds only (kimera-rpgo_b200/build.py::build_variant): an explicit path to another build of the
        # same library; still no fallback of any kind
        path = os.environ.get("RPGO_LIB_PATH", LIB_PATH)
        if path != LIB_PATH:
            lib = C.CDLL(path)
            for name, res, args in SYMBOLS:
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "kimera-rpgo_b200: %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
