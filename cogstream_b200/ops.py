"""Thin tensor-level wrappers over the C ABI (``include/cogstream_b200.h``).

PyTorch is plumbing here: it owns device memory and the current CUDA stream; every op passes
``data_ptr()``, sizes and the raw ``cudaStream_t`` to the shared library.  CUDA tensors only --
a CPU tensor, a missing library or a non-sm_100 device raises (no fallback of any kind).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import CSB_BF16, CSB_F32, CSB_F64, CompactArgs, KMeansParams, RelevanceArgs, STATUS_WORDS, check

_TS_DTYPE = {torch.float32: CSB_F32, torch.float64: CSB_F64}
_EL_DTYPE = {torch.float32: CSB_F32, torch.bfloat16: CSB_BF16}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need(t: torch.Tensor, name: str, dtype=None, dims: Optional[int] = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CPU tensor given; cogstream_b200 runs on sm_100a CUDA devices only (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if dims is not None and t.dim() != dims:
        raise ValueError(f"{name}: expected {dims} dims, got shape {tuple(t.shape)}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _on_device:
    """Make the tensor's device current for the duration of a call (cheap when it already is)."""

    def __init__(self, t: torch.Tensor):
        self.idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


def check_device() -> None:
    check(_lib.load().csb_check_device(), "check_device")


# --------------------------------------------------------------------------------------------- (a)
def pool_norm(x: torch.Tensor, *, want_bf16=True, want_f32=False, want_sqnorm=True):
    """x [T,P,D] bf16 -> (pooled_bf16 [T,D] | None, pooled_f32 [T,D] | None, sqnorm [T] fp32 | None)."""
    _need(x, "x", torch.bfloat16, 3)
    T, P, D = x.shape
    pb = torch.empty(T, D, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    pf = torch.empty(T, D, dtype=torch.float32, device=x.device) if want_f32 else None
    sq = torch.empty(T, dtype=torch.float32, device=x.device) if want_sqnorm else None
    with _on_device(x):
        check(_lib.load().csb_pool_norm(x.data_ptr(), T, P, D, _ptr(pb), _ptr(pf), _ptr(sq), _stream(x)), "pool_norm")
    return pb, pf, sq


# --------------------------------------------------------------------------------------------- (b)
def gram_workspace_bytes(T: int, K: int) -> int:
    return int(_lib.load().csb_gram_workspace_bytes(T, K))


def gram(x: torch.Tensor, *, impl: int = 0, out: Optional[torch.Tensor] = None,
         workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [T,K] bf16 -> G [T,T] fp64 (G = x x^T, fp32 tensor-core partials summed in fp64)."""
    _need(x, "x", torch.bfloat16, 2)
    T, K = x.shape
    G = out if out is not None else torch.empty(T, T, dtype=torch.float64, device=x.device)
    _need(G, "out", torch.float64, 2)
    need = gram_workspace_bytes(T, K)
    ws = workspace if workspace is not None and workspace.numel() * workspace.element_size() >= need else \
        torch.empty(need, dtype=torch.uint8, device=x.device)
    with _on_device(x):
        check(_lib.load().csb_gram(x.data_ptr(), T, K, G.data_ptr(), ws.data_ptr(), ws.numel() * ws.element_size(), impl,
                                   _stream(x)), "gram")
    return G


def _gram_view(G: torch.Tensor, name: str = "G"):
    """G may be the leading [T,T] corner of a larger row-major buffer: returns (T, leading dimension)."""
    if not (isinstance(G, torch.Tensor) and G.is_cuda and G.dtype == torch.float64 and G.dim() == 2 and G.shape[0] == G.shape[1]):
        raise TypeError(f"{name}: expected a square float64 CUDA matrix")
    if G.stride(1) != 1 or G.stride(0) < G.shape[1]:
        raise ValueError(f"{name}: rows must be contiguous (stride {G.stride()})")
    return G.shape[0], G.stride(0)


def gram_extend(x: torch.Tensor, T_old: int, G_buf: torch.Tensor, *, impl: int = 0, workspace: Optional[torch.Tensor] = None):
    """x [T,K] bf16 (ALL frames so far), G_buf [T_max,T_max] f64 whose [:T_old,:T_old] corner is valid:
    fills rows / columns [T_old, T) in place and returns the [T,T] view."""
    _need(x, "x", torch.bfloat16, 2)
    T, K = x.shape
    _need(G_buf, "G_buf", torch.float64, 2)
    ldg = G_buf.shape[1]
    if G_buf.shape[0] < T or ldg < T:
        raise ValueError("gram_extend: G_buf too small")
    need = gram_workspace_bytes(T, K)
    ws = workspace if workspace is not None and workspace.numel() >= need else torch.empty(need, dtype=torch.uint8, device=x.device)
    with _on_device(x):
        check(_lib.load().csb_gram_extend(x.data_ptr(), int(T_old), T, K, G_buf.data_ptr(), ldg, ws.data_ptr(), ws.numel(), impl,
                                          _stream(x)), "gram_extend")
    return G_buf[:T, :T]


def kmeans_workspace_bytes(T: int, k: int) -> int:
    return int(_lib.load().csb_kmeans_workspace_bytes(T, k))


class KMeansBuffers:
    """Caller-owned outputs + workspace of one csb_kmeans_solve call."""

    def __init__(self, T: int, k: int, ts_dtype: torch.dtype, device):
        self.T, self.k = T, k
        self.assignments = torch.empty(T, dtype=torch.int64, device=device)
        self.centroid_times = torch.empty(k, dtype=ts_dtype, device=device)
        self.member_offsets = torch.empty(k + 1, dtype=torch.int32, device=device)
        self.members = torch.empty(T + k, dtype=torch.int32, device=device)
        self.init_indices = torch.empty(k, dtype=torch.int64, device=device)
        self.event_order = torch.empty(T, dtype=torch.int32, device=device)  # frames sorted by (event, frame)
        self.event_offsets = torch.empty(k + 1, dtype=torch.int32, device=device)
        self.status = torch.zeros(STATUS_WORDS, dtype=torch.int32, device=device)
        self.workspace = torch.empty(kmeans_workspace_bytes(T, k), dtype=torch.uint8, device=device)

    def timeline(self):
        """Phase timeline of the last solve: list of (tag, ns since the first mark); profiling aid (synchronises)."""
        tail = self.workspace[-4096 + 256:].view(torch.int64).cpu()
        n = int(tail[0])
        ev = [(int(v) & 0xFF, int(v) >> 8) for v in tail[1: 1 + n].tolist()]
        t0 = ev[0][1] if ev else 0
        return [(tag, t - t0) for tag, t in ev]


def kmeans_solve(G: torch.Tensor, timestamps: torch.Tensor, buf: KMeansBuffers, *, n_init_given: int,
                 exp_noise: Optional[torch.Tensor], reseed_pool: Optional[torch.Tensor], alpha: float = 2.0,
                 max_iteration: int = 30, tol: float = 1e-4, frames_per_warp: int = 0,
                 rng_state: Optional[torch.Tensor] = None) -> None:
    """k-means++ + Lloyd on the Gram matrix.  ``buf.init_indices[:n_init_given]`` must be filled.
    ``rng_state`` (device int64[2] = {seed, offset}): the solver draws the Exp(1) noise, the re-seed indices and (with
    ``n_init_given == 0``) the first centre itself; ``exp_noise`` / ``reseed_pool`` may then be None."""
    _, ldg = _gram_view(G)
    _need(timestamps, "timestamps", None, 1)
    if timestamps.dtype not in _TS_DTYPE:
        raise TypeError(f"timestamps must be float32 or float64, got {timestamps.dtype}")
    T, k = buf.T, buf.k
    if G.shape != (T, T) or timestamps.shape[0] != T:
        raise ValueError("kmeans_solve: shape mismatch")
    if buf.centroid_times.dtype != timestamps.dtype:
        # the solver writes the centre times in the timestamps' dtype (kmeans_with_time.py:66): a float32 buffer would be overrun
        raise TypeError(f"kmeans_solve: buffers were allocated for {buf.centroid_times.dtype} timestamps, got {timestamps.dtype} "
                        "(KMeansBuffers / StreamCompressor ts_dtype)")
    if exp_noise is not None:
        _need(exp_noise, "exp_noise", torch.float32, 2)
        if exp_noise.shape != (k - 1, T):
            raise ValueError(f"exp_noise must be [k-1, T] = [{k - 1}, {T}], got {tuple(exp_noise.shape)}")
    n_pool = 0
    if reseed_pool is not None:
        _need(reseed_pool, "reseed_pool", torch.int64, 1)
        n_pool = reseed_pool.numel()
    if rng_state is not None:
        _need(rng_state, "rng_state", torch.int64, 1)
        if rng_state.numel() != 2:
            raise ValueError("rng_state must be an int64[2] tensor {seed, offset}")
    p = KMeansParams(T, k, _TS_DTYPE[timestamps.dtype], int(max_iteration), float(alpha), float(tol), int(n_init_given), n_pool, ldg,
                     _ptr(rng_state), int(frames_per_warp))
    with _on_device(G):
        check(_lib.load().csb_kmeans_solve(
            G.data_ptr(), timestamps.data_ptr(), C.byref(p), buf.init_indices.data_ptr(), _ptr(exp_noise), _ptr(reseed_pool),
            buf.assignments.data_ptr(), buf.centroid_times.data_ptr(), buf.member_offsets.data_ptr(), buf.members.data_ptr(),
            buf.event_order.data_ptr(), buf.event_offsets.data_ptr(), buf.status.data_ptr(), buf.workspace.data_ptr(), buf.workspace.numel(), _stream(G)), "kmeans_solve")


def decode_status(status_host: torch.Tensor) -> dict:
    """Decode a host copy of the int32[8] status block."""
    s = status_host.cpu()
    f = s.view(torch.float32)
    return dict(iterations=int(s[_lib.ST_ITER]), converged=bool(s[_lib.ST_CONV]), n_reseed_used=int(s[_lib.ST_RESEED]),
                degenerate_rounds=int(s[_lib.ST_DEGEN]), reseed_overflow=bool(s[_lib.ST_OVERFLOW]),
                n_noise_used=int(s[_lib.ST_NOISE]), min_margin=float(f[_lib.ST_MARGIN]), last_shift=float(f[_lib.ST_SHIFT]))


def centroid_update(x: torch.Tensor, member_offsets: torch.Tensor, members: torch.Tensor, k: int,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [T,K] bf16 + centre CSR -> centroids [k,K] fp32."""
    _need(x, "x", torch.bfloat16, 2)
    _need(member_offsets, "member_offsets", torch.int32, 1)
    _need(members, "members", torch.int32, 1)
    T, K = x.shape
    Cn = out if out is not None else torch.empty(k, K, dtype=torch.float32, device=x.device)
    with _on_device(x):
        check(_lib.load().csb_centroid_update(x.data_ptr(), T, K, member_offsets.data_ptr(), members.data_ptr(), k,
                                              Cn.data_ptr(), _stream(x)), "centroid_update")
    return Cn


def centroid_gather(x: torch.Tensor, member_offsets: torch.Tensor, members: torch.Tensor, event_offsets: torch.Tensor, k: int,
                    centroids: Optional[torch.Tensor] = None, packed: Optional[torch.Tensor] = None):
    """One pass over x [T,K] bf16: centroids [k,K] fp32 and/or the event-major packed copy [T,K] bf16 (outputs given by the caller)."""
    _need(x, "x", torch.bfloat16, 2)
    T, K = x.shape
    with _on_device(x):
        check(_lib.load().csb_centroid_gather(x.data_ptr(), T, K, member_offsets.data_ptr(), members.data_ptr(),
                                              event_offsets.data_ptr(), k, _ptr(centroids), _ptr(packed), _stream(x)),
              "centroid_gather")
    return centroids, packed


def dist_to_assigned(x: torch.Tensor, centroids: torch.Tensor, assignments: torch.Tensor) -> torch.Tensor:
    _need(x, "x", torch.bfloat16, 2)
    _need(centroids, "centroids", torch.float32, 2)
    _need(assignments, "assignments", torch.int64, 1)
    T, K = x.shape
    dist = torch.empty(T, dtype=torch.float32, device=x.device)
    with _on_device(x):
        check(_lib.load().csb_dist_to_assigned(x.data_ptr(), T, K, centroids.data_ptr(), centroids.shape[0],
                                               assignments.data_ptr(), dist.data_ptr(), _stream(x)), "dist_to_assigned")
    return dist


def dist_to_assigned_gram(G: torch.Tensor, assignments: torch.Tensor, event_offsets: torch.Tensor, order: torch.Tensor,
                          k: int) -> torch.Tensor:
    T, ldg = _gram_view(G)
    dist = torch.empty(T, dtype=torch.float32, device=G.device)
    ws = torch.empty(T, dtype=torch.float64, device=G.device)
    with _on_device(G):
        check(_lib.load().csb_dist_to_assigned_gram(G.data_ptr(), ldg, T, k, assignments.data_ptr(), event_offsets.data_ptr(),
                                                    order.data_ptr(), dist.data_ptr(), ws.data_ptr(), _stream(G)),
              "dist_to_assigned_gram")
    return dist


def select_representatives(dist: torch.Tensor, assignments: torch.Tensor, k: int, n_rep: int):
    """-> (reps [k,n_rep] int64 (-1 padded), rep_counts [k] int32, rep_flags [T] uint8)."""
    _need(dist, "dist", torch.float32, 1)
    _need(assignments, "assignments", torch.int64, 1)
    T = dist.shape[0]
    reps = torch.empty(k, n_rep, dtype=torch.int64, device=dist.device)
    counts = torch.empty(k, dtype=torch.int32, device=dist.device)
    flags = torch.empty(T, dtype=torch.uint8, device=dist.device)
    with _on_device(dist):
        check(_lib.load().csb_select_representatives(dist.data_ptr(), assignments.data_ptr(), T, k, n_rep, reps.data_ptr(),
                                                     counts.data_ptr(), flags.data_ptr(), _stream(dist)),
              "select_representatives")
    return reps, counts, flags


def representatives_gram(G: torch.Tensor, event_offsets: torch.Tensor, order: torch.Tensor, k: int, n_rep: int):
    """Fused dist_to_assigned_gram + select_representatives -> (dist [T] f32, reps [k,n_rep] i64, rep_counts [k] i32, rep_flags [T] u8)."""
    T, ldg = _gram_view(G)
    dev = G.device
    dist = torch.empty(T, dtype=torch.float32, device=dev)
    reps = torch.empty(k, n_rep, dtype=torch.int64, device=dev)
    counts = torch.empty(k, dtype=torch.int32, device=dev)
    flags = torch.empty(T, dtype=torch.uint8, device=dev)
    with _on_device(G):
        check(_lib.load().csb_representatives_gram(G.data_ptr(), ldg, T, k, event_offsets.data_ptr(), order.data_ptr(), n_rep,
                                                   dist.data_ptr(), reps.data_ptr(), counts.data_ptr(), flags.data_ptr(), _stream(G)),
              "representatives_gram")
    return dist, reps, counts, flags


def relevance_minor(question: torch.Tensor, events: torch.Tensor, assignments: torch.Tensor, rep_flags: torch.Tensor,
                    ticket: torch.Tensor, threshold: float = 0.45):
    """Fused cosine_flags + minor_frames -> (cos [k], low [k] u8, minor_flags [T] u8, minor_list [T] i32, n_minor [1] i64).
    ``ticket``: a zeroed int32[1] tensor owned by the caller (one per in-flight stream)."""
    _need(events, "events", None, 2)
    _need(question, "question", events.dtype, 1)
    if events.dtype not in _EL_DTYPE:
        raise TypeError(f"relevance_minor: dtype must be bfloat16 or float32, got {events.dtype}")
    k, Dl = events.shape
    T = assignments.shape[0]
    dev = events.device
    cos = torch.empty(k, dtype=events.dtype, device=dev)
    low = torch.empty(k, dtype=torch.uint8, device=dev)
    flags = torch.empty(T, dtype=torch.uint8, device=dev)
    lst = torch.empty(T, dtype=torch.int32, device=dev)
    n = torch.empty(1, dtype=torch.int64, device=dev)
    with _on_device(events):
        check(_lib.load().csb_relevance_minor(question.data_ptr(), events.data_ptr(), k, Dl, _EL_DTYPE[events.dtype], float(threshold),
                                              assignments.data_ptr(), rep_flags.data_ptr(), T, cos.data_ptr(), low.data_ptr(),
                                              flags.data_ptr(), lst.data_ptr(), n.data_ptr(), ticket.data_ptr(), _stream(events)),
              "relevance_minor")
    return cos, low, flags, lst, n


def event_order(assignments: torch.Tensor, k: int):
    """-> (order [T] int32: frames sorted by (event, frame), event_offsets [k+1] int32)."""
    _need(assignments, "assignments", torch.int64, 1)
    T = assignments.shape[0]
    order = torch.empty(T, dtype=torch.int32, device=assignments.device)
    offs = torch.empty(k + 1, dtype=torch.int32, device=assignments.device)
    with _on_device(assignments):
        check(_lib.load().csb_event_order(assignments.data_ptr(), T, k, order.data_ptr(), offs.data_ptr(),
                                          _stream(assignments)), "event_order")
    return order, offs


def gather_rows(src: torch.Tensor, index: torch.Tensor, out: torch.Tensor, *, n: Optional[int] = None,
                n_dev: Optional[torch.Tensor] = None, n_max: Optional[int] = None) -> torch.Tensor:
    """out[r] = src[index[r]] for r < n (host int) or r < n_dev[0] (device int64)."""
    _need(src, "src", None, 2)
    _need(index, "index", torch.int32, 1)
    _need(out, "out", src.dtype, 2)
    row_bytes = src.shape[1] * src.element_size()
    if n_max is None:
        n_max = n if n is not None else index.shape[0]
    with _on_device(src):
        check(_lib.load().csb_gather_rows(src.data_ptr(), row_bytes, index.data_ptr(), 0 if n is None else n, _ptr(n_dev),
                                          n_max, out.data_ptr(), _stream(src)), "gather_rows")
    return out


# --------------------------------------------------------------------------------------------- (c)
def cosine_flags(question: torch.Tensor, events: torch.Tensor, threshold: float = 0.45):
    """question [Dl] / events [k,Dl] (bf16 or fp32) -> (cos [k] same dtype, low_flags [k] uint8)."""
    _need(events, "events", None, 2)
    _need(question, "question", events.dtype, 1)
    if events.dtype not in _EL_DTYPE:
        raise TypeError(f"cosine_flags: dtype must be bfloat16 or float32, got {events.dtype}")
    k, Dl = events.shape
    cos = torch.empty(k, dtype=events.dtype, device=events.device)
    low = torch.empty(k, dtype=torch.uint8, device=events.device)
    with _on_device(events):
        check(_lib.load().csb_cosine_flags(question.data_ptr(), events.data_ptr(), k, Dl, _EL_DTYPE[events.dtype],
                                           float(threshold), cos.data_ptr(), low.data_ptr(), _stream(events)), "cosine_flags")
    return cos, low


def minor_frames(assignments: torch.Tensor, low_flags: torch.Tensor, rep_flags: torch.Tensor):
    """-> (minor_flags [T] uint8, minor_list [T] int32 (first n_minor valid, ascending), n_minor [1] int64)."""
    T = assignments.shape[0]
    dev = assignments.device
    flags = torch.empty(T, dtype=torch.uint8, device=dev)
    lst = torch.empty(T, dtype=torch.int32, device=dev)
    n = torch.empty(1, dtype=torch.int64, device=dev)
    with _on_device(assignments):
        check(_lib.load().csb_minor_frames(assignments.data_ptr(), low_flags.data_ptr(), rep_flags.data_ptr(), T,
                                           flags.data_ptr(), lst.data_ptr(), n.data_ptr(), _stream(assignments)), "minor_frames")
    return flags, lst, n


# --------------------------------------------------------------------------------------------- (d)
def pixel_mask(pixels: torch.Tensor, t: int, P: int, W: int, *, threshold: float = 0.1, min_tokens: int = 1,
               minor_flags: Optional[torch.Tensor] = None, frame0: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One video item: pixels (t*P*W elements, bf16 or fp32) -> mask [t*P] uint8."""
    _need(pixels, "pixels")
    if pixels.dtype not in _EL_DTYPE:
        raise TypeError(f"pixel_mask: dtype must be bfloat16 or float32, got {pixels.dtype}")
    if pixels.numel() != t * P * W:
        raise ValueError("pixel_mask: pixel count does not match t*P*W")
    mask = out if out is not None else torch.empty(t * P, dtype=torch.uint8, device=pixels.device)
    with _on_device(pixels):
        check(_lib.load().csb_pixel_mask(pixels.data_ptr(), t, P, W, _EL_DTYPE[pixels.dtype], float(threshold), int(min_tokens),
                                         _ptr(minor_flags), int(frame0), mask.data_ptr(), _stream(pixels)), "pixel_mask")
    return mask


def pixel_diff(pixels: torch.Tensor, t: int, P: int, W: int, *, threshold: float = 0.1, out: Optional[torch.Tensor] = None,
               stream: Optional[torch.cuda.Stream] = None, max_ctas: int = 0, evict_first: bool = True) -> torch.Tensor:
    """Raw pixel-difference mask [t*P] uint8 (no per-frame rules); ``stream`` overrides the current stream.  ``max_ctas`` > 0
    throttles the grid (for running next to the latency-bound k-means solver), ``evict_first`` = L2 evict-first pixel loads."""
    _need(pixels, "pixels")
    if pixels.dtype not in _EL_DTYPE:
        raise TypeError(f"pixel_diff: dtype must be bfloat16 or float32, got {pixels.dtype}")
    if pixels.numel() != t * P * W:
        raise ValueError("pixel_diff: pixel count does not match t*P*W")
    mask = out if out is not None else torch.empty(t * P, dtype=torch.uint8, device=pixels.device)
    st = stream.cuda_stream if stream is not None else _stream(pixels)
    with _on_device(pixels):
        check(_lib.load().csb_pixel_diff_ex(pixels.data_ptr(), t, P, W, _EL_DTYPE[pixels.dtype], float(threshold), mask.data_ptr(),
                                            int(max_ctas), 1 if evict_first else 0, st), "pixel_diff")
    return mask


def mask_finalize(mask: torch.Tensor, t: int, P: int, *, min_tokens: int = 1, minor_flags: Optional[torch.Tensor] = None,
                  frame0: int = 0) -> torch.Tensor:
    """In place: min_tokens padding, then the minor-frame override (cogreasoner_chat.py:413-422)."""
    _need(mask, "mask", torch.uint8, 1)
    with _on_device(mask):
        check(_lib.load().csb_mask_finalize(mask.data_ptr(), t, P, int(min_tokens), _ptr(minor_flags), int(frame0), _stream(mask)),
              "mask_finalize")
    return mask


def pool_frames_inplace(x: torch.Tensor, P: int, frames: torch.Tensor, *, n: Optional[int] = None,
                        n_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [T*P, D] bf16 (a private clone): token 0 of every listed frame <- mean of the frame's tokens."""
    _need(x, "x", torch.bfloat16, 2)
    _need(frames, "frames", torch.int32, 1)
    TP, D = x.shape
    T = TP // P
    with _on_device(x):
        check(_lib.load().csb_pool_frames_inplace(x.data_ptr(), T, P, D, frames.data_ptr(),
                                                  frames.shape[0] if n is None else n, _ptr(n_dev), _stream(x)),
              "pool_frames_inplace")
    return x


def compact_workspace_bytes(n_tokens: int, n_ids: int) -> int:
    return int(_lib.load().csb_compact_workspace_bytes(n_tokens, n_ids))


def compact_tokens(x: torch.Tensor, T: int, P: int, mask: torch.Tensor, pool_flags: Optional[torch.Tensor] = None, *,
                   out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None):
    """x [T*P, D] bf16, mask [T*P] uint8 -> (out [T*P, D] (first n_kept rows valid), kept_index [T*P] int32, n_kept [1] int64)."""
    _need(x, "x", torch.bfloat16, 2)
    _need(mask, "mask", torch.uint8, 1)
    D = x.shape[1]
    n = T * P
    dev = x.device
    o = out if out is not None else torch.empty(n, D, dtype=torch.bfloat16, device=dev)
    kept = torch.empty(n, dtype=torch.int32, device=dev)
    n_kept = torch.empty(1, dtype=torch.int64, device=dev)
    need = compact_workspace_bytes(n, 0)
    ws = workspace if workspace is not None and workspace.numel() >= need else torch.empty(need, dtype=torch.uint8, device=dev)
    with _on_device(x):
        check(_lib.load().csb_compact_tokens(x.data_ptr(), T, P, D, mask.data_ptr(), _ptr(pool_flags), o.data_ptr(),
                                             kept.data_ptr(), n_kept.data_ptr(), ws.data_ptr(), ws.numel(), _stream(x)),
              "compact_tokens")
    return o, kept, n_kept


def filter_ids(input_ids: torch.Tensor, image_token_index: int, mask: torch.Tensor,
               attention_mask: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
               position_ids: Optional[torch.Tensor] = None):
    """-> (out_ids, out_attention_mask, out_labels, out_position_ids, n_out [2] int64 = {ids kept, <image> ids seen}); outputs
    sized like the inputs.  The caller compares ``n_out[1]`` with ``mask.numel()`` (the reference raises when they differ)."""
    _need(input_ids, "input_ids", torch.int64, 1)
    _need(mask, "mask", torch.uint8, 1)
    n = input_ids.shape[0]
    dev = input_ids.device

    def _opt(t, name):
        if t is None:
            return None, None
        if t.dtype != torch.int64:
            t = t.to(torch.int64)
        _need(t, name, torch.int64, 1)
        return t, torch.empty_like(t)

    am, o_am = _opt(attention_mask, "attention_mask")
    lb, o_lb = _opt(labels, "labels")
    ps, o_ps = _opt(position_ids, "position_ids")
    o_ids = torch.empty_like(input_ids)
    n_out = torch.empty(2, dtype=torch.int64, device=dev)
    ws = torch.empty(compact_workspace_bytes(0, n), dtype=torch.uint8, device=dev)
    with _on_device(input_ids):
        check(_lib.load().csb_filter_ids(input_ids.data_ptr(), n, int(image_token_index), mask.data_ptr(), mask.shape[0],
                                         _ptr(am), _ptr(lb), _ptr(ps), o_ids.data_ptr(), _ptr(o_am), _ptr(o_lb), _ptr(o_ps),
                                         n_out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(input_ids)), "filter_ids")
    return o_ids, o_am, o_lb, o_ps, n_out


def compact_stream_workspace(n_tokens: int, n_ids: int, device) -> torch.Tensor:
    """Zero-filled workspace of :func:`compact_stream` (it carries the barrier epoch from launch to launch: one per CUDA stream)."""
    return torch.zeros(int(_lib.load().csb_compact_stream_workspace_bytes(n_tokens, n_ids)), dtype=torch.uint8, device=device)


def compact_stream(x: torch.Tensor, T: int, P: int, mask: torch.Tensor, minor_flags: Optional[torch.Tensor], *, min_tokens: int = 1,
                   finalize: bool = True, pool_minor: bool = True, input_ids: Optional[torch.Tensor] = None,
                   image_token_index: int = -1, attention_mask: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                   out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None, bufs: Optional[dict] = None,
                   frame_kind: Optional[torch.Tensor] = None):
    """Phase-2 tail in one entry point: per-frame mask rules (in place on ``mask``), token compaction with the pooled token of the
    minor frames, id filtering.  -> dict(out [T*P, D], kept_index [T*P] i32, n_kept [1] i64, out_ids, out_attention, out_labels,
    n_out [2] i64 = {ids kept, <image> ids seen}).  ``bufs``: a dict of pre-allocated outputs to reuse (same keys).
    ``frame_kind`` [T] uint8: 1 marks the frames of image items / single-frame videos (all tokens kept, no per-frame rule)."""
    _need(x, "x", torch.bfloat16, 2)
    _need(mask, "mask", torch.uint8, 1)
    D = x.shape[1]
    n = T * P
    dev = x.device
    b = bufs if bufs is not None else {}

    def buf(name, shape, dtype):
        t = b.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=dev)
            b[name] = t
        return t

    o = out if out is not None else buf("out", (n, D), torch.bfloat16)
    b["out"] = o
    kept = buf("kept_index", (n,), torch.int32)
    n_kept = buf("n_kept", (1,), torch.int64)
    n_ids = 0
    o_ids = o_am = o_lb = n_out = None
    if input_ids is not None:
        _need(input_ids, "input_ids", torch.int64, 1)
        n_ids = input_ids.shape[0]
        o_ids = buf("out_ids", (n_ids,), torch.int64)
        n_out = buf("n_out", (2,), torch.int64)
        if attention_mask is not None:
            if attention_mask.dtype != torch.int64:
                attention_mask = attention_mask.to(torch.int64)
            o_am = buf("out_attention", (n_ids,), torch.int64)
        if labels is not None:
            if labels.dtype != torch.int64:
                labels = labels.to(torch.int64)
            o_lb = buf("out_labels", (n_ids,), torch.int64)
    need = int(_lib.load().csb_compact_stream_workspace_bytes(n, n_ids))
    ws = workspace if workspace is not None and workspace.numel() >= need else torch.zeros(need, dtype=torch.uint8, device=dev)
    a = CompactArgs(x.data_ptr(), T, P, D, mask.data_ptr(), _ptr(minor_flags), _ptr(frame_kind), int(min_tokens), 1 if finalize else 0,
                    1 if (pool_minor and minor_flags is not None) else 0, _ptr(input_ids), n_ids, int(image_token_index),
                    _ptr(attention_mask), _ptr(labels), o.data_ptr(), kept.data_ptr(), n_kept.data_ptr(), _ptr(o_ids), _ptr(o_am),
                    _ptr(o_lb), _ptr(n_out), ws.data_ptr(), ws.numel())
    with _on_device(x):
        check(_lib.load().csb_compact_stream(C.byref(a), _stream(x)), "compact_stream")
    return b


def relevance_select(question: torch.Tensor, events: Optional[torch.Tensor], G: torch.Tensor, event_offsets: torch.Tensor,
                     event_order: torch.Tensor, assignments: torch.Tensor, ticket: torch.Tensor, *, k: int, n_rep: int = 2,
                     threshold: float = 0.45, hidden: Optional[torch.Tensor] = None, cu_seqlens: Optional[torch.Tensor] = None,
                     bufs: Optional[dict] = None):
    """Selection step of phase 2 in one launch: cosine flags of the k events, representatives from the Gram matrix, minor-frame
    list.  Event embeddings: ``events`` [k, Dl], or the packed ``hidden`` states + ``cu_seqlens`` of the k summary forwards (their
    mean pool is then fused in; the pooled rows are returned as ``events``).
    -> dict(cos, low, dist, reps, rep_counts, rep_flags, minor_flags, minor_list, n_minor[, events])."""
    T, ldg = _gram_view(G)
    src = events if events is not None else hidden
    _need(src, "events / hidden", None, 2)
    if src.dtype not in _EL_DTYPE:
        raise TypeError(f"relevance_select: dtype must be bfloat16 or float32, got {src.dtype}")
    _need(question, "question", src.dtype, 1)
    Dl = src.shape[1]
    dev = src.device
    b = bufs if bufs is not None else {}

    def buf(name, shape, dtype):
        t = b.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=dev)
            b[name] = t
        return t

    cos = buf("cos", (k,), src.dtype)
    low = buf("low", (k,), torch.uint8)
    dist = buf("dist", (T,), torch.float32)
    reps = buf("reps", (k, n_rep), torch.int64)
    counts = buf("rep_counts", (k,), torch.int32)
    rflags = buf("rep_flags", (T,), torch.uint8)
    mflags = buf("minor_flags", (T,), torch.uint8)
    mlist = buf("minor_list", (T,), torch.int32)
    n_minor = buf("n_minor", (1,), torch.int64)
    ev_out = ws = None
    if events is None:
        _need(cu_seqlens, "cu_seqlens", torch.int32, 1)
        if cu_seqlens.shape[0] != k + 1:
            raise ValueError("relevance_select: cu_seqlens must have k + 1 entries")
        ev_out = buf("events", (k, Dl), src.dtype)
        ws = buf("pool_ws", (int(_lib.load().csb_relevance_select_workspace_bytes(k, Dl)),), torch.uint8)
    elif events.shape[0] != k:
        raise ValueError("relevance_select: events must have k rows")
    a = RelevanceArgs(question.data_ptr(), _ptr(events), _ptr(hidden) if events is None else None, _ptr(cu_seqlens) if events is None else None,
                      k, Dl, _EL_DTYPE[src.dtype], 0, float(threshold), G.data_ptr(), ldg, event_offsets.data_ptr(), event_order.data_ptr(),
                      int(n_rep), assignments.data_ptr(), T, _ptr(ev_out), cos.data_ptr(), low.data_ptr(), dist.data_ptr(), reps.data_ptr(),
                      counts.data_ptr(), rflags.data_ptr(), mflags.data_ptr(), mlist.data_ptr(), n_minor.data_ptr(), ticket.data_ptr(),
                      _ptr(ws), 0 if ws is None else ws.numel())
    with _on_device(src):
        check(_lib.load().csb_relevance_select(C.byref(a), _stream(src)), "relevance_select")
    return b


def embed_scatter(input_ids: torch.Tensor, image_token_index: int, embed_weight: torch.Tensor, features: Optional[torch.Tensor],
                  *, out: Optional[torch.Tensor] = None, check_count: bool = True, frame_order: Optional[torch.Tensor] = None,
                  tokens_per_frame: int = 1):
    """``embed_tokens(input_ids)`` with the rows of the <image> ids replaced by ``embeds * 0.0 + features``
    (cogreasoner_chat.py:303-305, :567-572) in one pass.  ``input_ids`` [n] int64 may be several prompts packed back to back;
    ``features`` [n_image, D] are consumed in order.  -> (inputs_embeds [n, D], status [2] int32 = {n_image, n_bad_ids}).

    ``frame_order`` [T] int32 + ``tokens_per_frame``: read the feature rows through the chronological per-event frame order
    (the j-th <image> id takes row ``frame_order[j // P] * P + j % P``) -- the event gather of :286-296 folded into this pass.

    ``check_count=True`` reads the status back (one sync) and raises like the reference's index assignment would when the
    number of <image> ids differs from the number of feature rows, or an id is outside the table."""
    _need(input_ids, "input_ids", torch.int64, 1)
    _need(embed_weight, "embed_weight", None, 2)
    if embed_weight.dtype not in _EL_DTYPE:
        raise TypeError("embed_scatter: embedding table must be bf16 or f32")
    n, (vocab, D) = input_ids.shape[0], embed_weight.shape
    dev = input_ids.device
    if features is None:
        features = torch.empty(0, D, dtype=embed_weight.dtype, device=dev)
    if features.dtype != embed_weight.dtype:  # `.to(inputs_embeds.dtype)` of the reference
        features = features.to(embed_weight.dtype)
    _need(features, "features", embed_weight.dtype, 2)
    if features.shape[1] != D:
        raise ValueError(f"embed_scatter: feature width {features.shape[1]} != embedding width {D}")
    o = out if out is not None else torch.empty(n, D, dtype=embed_weight.dtype, device=dev)
    status = torch.empty(2, dtype=torch.int32, device=dev)
    ws = torch.empty(int(_lib.load().csb_embed_scatter_workspace_bytes(n)), dtype=torch.uint8, device=dev)
    with _on_device(input_ids):
        if frame_order is not None:
            _need(frame_order, "frame_order", torch.int32, 1)
            if frame_order.shape[0] * int(tokens_per_frame) != features.shape[0]:
                raise ValueError("embed_scatter: frame_order x tokens_per_frame does not cover the feature rows")
            check(_lib.load().csb_embed_scatter_ordered(input_ids.data_ptr(), n, int(image_token_index), embed_weight.data_ptr(), vocab,
                                                        features.data_ptr(), frame_order.data_ptr(), frame_order.shape[0],
                                                        int(tokens_per_frame), D, _EL_DTYPE[embed_weight.dtype], o.data_ptr(),
                                                        status.data_ptr(), ws.data_ptr(), ws.numel(), _stream(input_ids)),
                  "embed_scatter_ordered")
        else:
            check(_lib.load().csb_embed_scatter(input_ids.data_ptr(), n, int(image_token_index), embed_weight.data_ptr(), vocab,
                                                features.data_ptr() if features.shape[0] else None, features.shape[0], D,
                                                _EL_DTYPE[embed_weight.dtype], o.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                                ws.numel(), _stream(input_ids)), "embed_scatter")
    if check_count:
        n_img, n_bad = status.tolist()
        if n_bad:
            raise IndexError(f"embed_scatter: {n_bad} input ids outside the embedding table [0, {vocab})")
        if n_img != features.shape[0]:
            raise RuntimeError(f"shape mismatch: value tensor of shape [{features.shape[0]}, {D}] cannot be broadcast to "
                               f"indexing result of shape [{n_img}, {D}]")
    return o, status


def seq_mean_pool(hidden: torch.Tensor, cu_seqlens: torch.Tensor, *, out: Optional[torch.Tensor] = None):
    """``torch.mean(last_hidden_state, dim=1)`` of every sequence of a packed var-len batch (cogreasoner_chat.py:317-318, :322):
    ``hidden`` [sum(len), Dl] bf16 / f32, ``cu_seqlens`` [n_seq + 1] int32 -> [n_seq, Dl]."""
    _need(hidden, "hidden", None, 2)
    if hidden.dtype not in _EL_DTYPE:
        raise TypeError("seq_mean_pool: hidden must be bf16 or f32")
    _need(cu_seqlens, "cu_seqlens", torch.int32, 1)
    n_seq, Dl = cu_seqlens.shape[0] - 1, hidden.shape[1]
    dev = hidden.device
    o = out if out is not None else torch.empty(n_seq, Dl, dtype=hidden.dtype, device=dev)
    ws = torch.empty(int(_lib.load().csb_seq_mean_pool_workspace_bytes(n_seq, Dl)), dtype=torch.uint8, device=dev)
    with _on_device(hidden):
        check(_lib.load().csb_seq_mean_pool(hidden.data_ptr(), cu_seqlens.data_ptr(), n_seq, Dl, _EL_DTYPE[hidden.dtype],
                                            o.data_ptr(), ws.data_ptr(), ws.numel(), _stream(hidden)), "seq_mean_pool")
    return o


# --------------------------------------------------------------------------------------------- upstream: encoder tail
def ln_merge(hidden: torch.Tensor, merge: int, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``post_layernorm`` + the ``merge x merge`` bilinear token merge of the vision tower in one pass
    (modeling_videollama3_encoder.py:479, :485-502).  ``hidden`` [n, C] bf16, merge-block-contiguous -> [n / merge^2, C] bf16."""
    _need(hidden, "hidden", torch.bfloat16, 2)
    _need(gamma, "gamma", torch.bfloat16, 1)
    _need(beta, "beta", torch.bfloat16, 1)
    n, Cw = hidden.shape
    if gamma.shape[0] != Cw or beta.shape[0] != Cw:
        raise ValueError("ln_merge: LayerNorm weight / bias width mismatch")
    m2 = int(merge) * int(merge)
    if n % m2:
        raise ValueError(f"ln_merge: {n} rows are not whole {merge} x {merge} blocks")
    o = out if out is not None else torch.empty(n // m2, Cw, dtype=torch.bfloat16, device=hidden.device)
    with _on_device(hidden):
        check(_lib.load().csb_ln_merge(hidden.data_ptr(), n, Cw, int(merge), gamma.data_ptr(), beta.data_ptr(), float(eps), o.data_ptr(),
                                       _stream(hidden)), "ln_merge")
    return o
