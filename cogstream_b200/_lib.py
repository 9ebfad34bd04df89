"""ctypes binding of ``include/cogstream_b200.h`` (the C-ABI shared library).

The library is built in-tree (``cogstream_b200/lib/libcogstream_b200.so``) by
``__graft_entry__.build()`` / ``make -C cogstream_b200/csrc``.  There is NO fallback: if the
library is missing, or the device is not sm_100, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcogstream_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

CSB_F32, CSB_F64, CSB_BF16 = 0, 1, 2


class KMeansParams(C.Structure):
    _fields_ = [
        ("T", C.c_int64),
        ("k", C.c_int64),
        ("ts_dtype", C.c_int32),
        ("max_iteration", C.c_int32),
        ("alpha", C.c_double),
        ("tol", C.c_double),
        ("n_init_given", C.c_int64),
        ("n_reseed_pool", C.c_int64),
        ("ldg", C.c_int64),
        ("rng_state", C.c_void_p),
        ("frames_per_warp", C.c_int64),
    ]


class CompactArgs(C.Structure):
    """csb_compact_args (include/cogstream_b200.h)."""
    _fields_ = [
        ("x_bf16", C.c_void_p), ("T", C.c_int64), ("P", C.c_int64), ("D", C.c_int64),
        ("mask", C.c_void_p), ("minor_flags", C.c_void_p), ("frame_kind", C.c_void_p), ("min_tokens", C.c_int64),
        ("finalize", C.c_int32), ("pool_minor", C.c_int32),
        ("input_ids", C.c_void_p), ("n_ids", C.c_int64), ("image_token_index", C.c_int64),
        ("attention_mask", C.c_void_p), ("labels", C.c_void_p),
        ("out_bf16", C.c_void_p), ("kept_index", C.c_void_p), ("n_kept", C.c_void_p),
        ("out_ids", C.c_void_p), ("out_attention_mask", C.c_void_p), ("out_labels", C.c_void_p), ("n_out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class RelevanceArgs(C.Structure):
    """csb_relevance_args (include/cogstream_b200.h)."""
    _fields_ = [
        ("question", C.c_void_p), ("events", C.c_void_p), ("hidden", C.c_void_p), ("cu_seqlens", C.c_void_p),
        ("k", C.c_int64), ("Dl", C.c_int64), ("dtype", C.c_int32), ("reserved", C.c_int32), ("threshold", C.c_double),
        ("gram", C.c_void_p), ("ldg", C.c_int64), ("event_offsets", C.c_void_p), ("event_order", C.c_void_p), ("n_rep", C.c_int64),
        ("assignments", C.c_void_p), ("T", C.c_int64),
        ("events_out", C.c_void_p), ("cos_out", C.c_void_p), ("low_flags", C.c_void_p), ("dist", C.c_void_p), ("reps", C.c_void_p),
        ("rep_counts", C.c_void_p), ("rep_flags", C.c_void_p), ("minor_flags", C.c_void_p), ("minor_list", C.c_void_p),
        ("n_minor", C.c_void_p), ("ticket", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class LloydShard(C.Structure):
    """csb_lloyd_shard (include/cogstream_b200.h)."""
    _fields_ = [("phase", C.c_int32), ("reserved", C.c_int32), ("f_begin", C.c_int64), ("f_end", C.c_int64), ("per_rank", C.c_int64),
                ("exchange", C.c_void_p), ("state", C.c_void_p), ("state_bytes", C.c_size_t)]


LLOYD_INIT, LLOYD_ASSIGN, LLOYD_UPDATE, LLOYD_SHIFT, LLOYD_OUTPUT = 1, 2, 3, 4, 5

# layout of the device-side csb_kmeans_status (8 x 4 bytes); read back as an int32[8] tensor
STATUS_WORDS = 8
ST_ITER, ST_CONV, ST_RESEED, ST_DEGEN, ST_OVERFLOW, ST_NOISE, ST_MARGIN, ST_SHIFT = range(8)

_p, _i64, _i32, _dbl, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/cogstream_b200.h one to one
PROTOTYPES = {
    "csb_abi_version": (C.c_int, []),
    "csb_last_error": (C.c_char_p, []),
    "csb_check_device": (C.c_int, []),
    "csb_kernel_launches": (C.c_uint64, []),
    "csb_pool_norm": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "csb_gram_workspace_bytes": (_sz, [_i64, _i64]),
    "csb_gram": (C.c_int, [_p, _i64, _i64, _p, _p, _sz, C.c_int, _p]),
    "csb_gram_extend": (C.c_int, [_p, _i64, _i64, _i64, _p, _i64, _p, _sz, C.c_int, _p]),
    "csb_kmeans_workspace_bytes": (_sz, [_i64, _i64]),
    "csb_kmeans_solve": (C.c_int, [_p, _p, C.POINTER(KMeansParams), _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "csb_kmeans_lloyd_state_bytes": (_sz, [_i64, _i64]),
    "csb_kmeans_lloyd_phase": (C.c_int, [_p, _p, C.POINTER(KMeansParams), C.POINTER(LloydShard), _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "csb_centroid_update": (C.c_int, [_p, _i64, _i64, _p, _p, _i64, _p, _p]),
    "csb_centroid_gather": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _i64, _p, _p, _p]),
    "csb_dist_to_assigned": (C.c_int, [_p, _i64, _i64, _p, _i64, _p, _p, _p]),
    "csb_dist_to_assigned_gram": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "csb_select_representatives": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "csb_representatives_gram": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _i64, _p, _p, _p, _p, _p]),
    "csb_relevance_minor": (C.c_int, [_p, _p, _i64, _i64, _i32, _dbl, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "csb_event_order": (C.c_int, [_p, _i64, _i64, _p, _p, _p]),
    "csb_gather_rows": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _p, _p]),
    "csb_cosine_flags": (C.c_int, [_p, _p, _i64, _i64, _i32, _dbl, _p, _p, _p]),
    "csb_minor_frames": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p, _p]),
    "csb_pixel_diff": (C.c_int, [_p, _i64, _i64, _i64, _i32, _dbl, _p, _p]),
    "csb_pixel_diff_ex": (C.c_int, [_p, _i64, _i64, _i64, _i32, _dbl, _p, _i64, _i32, _p]),
    "csb_mask_finalize": (C.c_int, [_p, _i64, _i64, _i64, _p, _i64, _p]),
    "csb_pixel_mask": (C.c_int, [_p, _i64, _i64, _i64, _i32, _dbl, _i64, _p, _i64, _p, _p]),
    "csb_pool_frames_inplace": (C.c_int, [_p, _i64, _i64, _i64, _p, _i64, _p, _p]),
    "csb_compact_workspace_bytes": (_sz, [_i64, _i64]),
    "csb_compact_tokens": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "csb_filter_ids": (C.c_int, [_p, _i64, _i64, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "csb_compact_stream_workspace_bytes": (_sz, [_i64, _i64]),
    "csb_compact_stream": (C.c_int, [C.POINTER(CompactArgs), _p]),
    "csb_relevance_select_workspace_bytes": (_sz, [_i64, _i64]),
    "csb_relevance_select": (C.c_int, [C.POINTER(RelevanceArgs), _p]),
    "csb_embed_scatter_workspace_bytes": (_sz, [_i64]),
    "csb_embed_scatter_ordered": (C.c_int, [_p, _i64, _i64, _p, _i64, _p, _p, _i64, _i64, _i64, _i32, _p, _p, _p, _sz, _p]),
    "csb_embed_scatter": (C.c_int, [_p, _i64, _i64, _p, _i64, _p, _i64, _i64, _i32, _p, _p, _p, _sz, _p]),
    "csb_ln_merge": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _dbl, _p, _p]),
    "csb_seq_mean_pool_workspace_bytes": (_sz, [_i64, _i64]),
    "csb_seq_mean_pool": (C.c_int, [_p, _p, _i64, _i64, _i32, _p, _p, _sz, _p]),
}

_lib: Optional[C.CDLL] = None


class CogStreamLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CogStreamLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C {CSRC_DIR}`. cogstream_b200 has no CPU / PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().csb_last_error().decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    if status != 0:
        raise RuntimeError(f"cogstream_b200.{what} failed ({status}): {last_error()}")
