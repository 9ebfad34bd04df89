"""torch.library registration of the C-ABI kernels (namespace ``cogstream_b200``): every compute entry point of
``include/cogstream_b200.h`` is reachable as ``torch.ops.cogstream_b200.<name>``.

The ops are thin: each implementation forwards the tensors' ``data_ptr()`` / sizes / the current stream to the shared
library through :mod:`cogstream_b200.ops` (ctypes).  CUDA only (``device_types="cuda"``): calling one with CPU tensors
raises ``NotImplementedError`` from the dispatcher -- there is no CPU kernel to fall back to.  Fake (meta)
implementations give shapes / dtypes so the ops compose with ``torch.compile`` / export without running anything.
Ops that fill caller-owned buffers (``kmeans_solve``, ``pool_frames_inplace``, ``mask_finalize``) declare them in
``mutates_args``.  The fused pipeline (:mod:`cogstream_b200.pipeline`) calls :mod:`cogstream_b200.ops` directly: same
kernels, minus the dispatcher's per-call host overhead.

    gram(x, impl) -> G                                   gram_extend(x, T_old, G_buf, impl) -> ()
    kmeans_solve(G, ts, init, noise?, pool?, ...)        centroid_update / centroid_gather
    dist_to_assigned(_gram), select_representatives, representatives_gram, event_order, gather_rows
    cosine_flags, minor_frames, relevance_minor, relevance_select(_hidden)
    pixel_diff, mask_finalize, pixel_mask, pool_norm, pool_frames_inplace
    compact_tokens, filter_ids, compact_stream, embed_scatter(_ordered), seq_mean_pool
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.library import custom_op

from . import ops

NS = "cogstream_b200"
T_ = torch.Tensor


def _op(name, mutates=()):
    return custom_op(f"{NS}::{name}", mutates_args=mutates, device_types="cuda")


# ------------------------------------------------------------------------------------------------ (b) k-means
@_op("gram")
def gram(x: T_, impl: int = 0) -> T_:
    return ops.gram(x, impl=impl)


@gram.register_fake
def _(x, impl=0):
    return x.new_empty((x.shape[0], x.shape[0]), dtype=torch.float64)


@_op("gram_extend", mutates=("G_buf",))
def gram_extend(x: T_, T_old: int, G_buf: T_, impl: int = 0) -> None:
    ops.gram_extend(x, T_old, G_buf, impl=impl)


@gram_extend.register_fake
def _(x, T_old, G_buf, impl=0):
    return None


@_op("kmeans_solve", mutates=("init_indices", "assignments", "centroid_times", "member_offsets", "members", "event_order",
                              "event_offsets", "status", "workspace"))
def kmeans_solve(G: T_, timestamps: T_, init_indices: T_, exp_noise: Optional[T_], reseed_pool: Optional[T_], assignments: T_,
                 centroid_times: T_, member_offsets: T_, members: T_, event_order: T_, event_offsets: T_, status: T_, workspace: T_,
                 n_init_given: int, alpha: float = 2.0, max_iteration: int = 30, tol: float = 1e-4) -> None:
    """k-means++ + Lloyd on the Gram matrix into caller-owned buffers (the fields of ``ops.KMeansBuffers``)."""
    buf = ops.KMeansBuffers.__new__(ops.KMeansBuffers)
    buf.T, buf.k = assignments.shape[0], centroid_times.shape[0]
    buf.assignments, buf.centroid_times, buf.member_offsets, buf.members = assignments, centroid_times, member_offsets, members
    buf.init_indices, buf.event_order, buf.event_offsets, buf.status, buf.workspace = init_indices, event_order, event_offsets, status, workspace
    ops.kmeans_solve(G, timestamps, buf, n_init_given=n_init_given, exp_noise=exp_noise, reseed_pool=reseed_pool, alpha=alpha,
                     max_iteration=max_iteration, tol=tol)


@kmeans_solve.register_fake
def _(*a, **k):
    return None


@_op("centroid_update")
def centroid_update(x: T_, member_offsets: T_, members: T_, k: int) -> T_:
    return ops.centroid_update(x, member_offsets, members, k)


@centroid_update.register_fake
def _(x, member_offsets, members, k):
    return x.new_empty((k, x.shape[1]), dtype=torch.float32)


@_op("centroid_gather")
def centroid_gather(x: T_, member_offsets: T_, members: T_, event_offsets: T_, k: int) -> Tuple[T_, T_]:
    """-> (centroids [k,K] f32, packed event-major copy [T,K] bf16)."""
    c = torch.empty(k, x.shape[1], dtype=torch.float32, device=x.device)
    p = torch.empty_like(x)
    ops.centroid_gather(x, member_offsets, members, event_offsets, k, c, p)
    return c, p


@centroid_gather.register_fake
def _(x, member_offsets, members, event_offsets, k):
    return x.new_empty((k, x.shape[1]), dtype=torch.float32), torch.empty_like(x)


@_op("dist_to_assigned")
def dist_to_assigned(x: T_, centroids: T_, assignments: T_) -> T_:
    return ops.dist_to_assigned(x, centroids, assignments)


@dist_to_assigned.register_fake
def _(x, centroids, assignments):
    return x.new_empty((x.shape[0],), dtype=torch.float32)


@_op("dist_to_assigned_gram")
def dist_to_assigned_gram(G: T_, assignments: T_, event_offsets: T_, order: T_, k: int) -> T_:
    return ops.dist_to_assigned_gram(G, assignments, event_offsets, order, k)


@dist_to_assigned_gram.register_fake
def _(G, assignments, event_offsets, order, k):
    return G.new_empty((G.shape[0],), dtype=torch.float32)


@_op("select_representatives")
def select_representatives(dist: T_, assignments: T_, k: int, n_rep: int) -> Tuple[T_, T_, T_]:
    return ops.select_representatives(dist, assignments, k, n_rep)


@select_representatives.register_fake
def _(dist, assignments, k, n_rep):
    return (dist.new_empty((k, n_rep), dtype=torch.int64), dist.new_empty((k,), dtype=torch.int32),
            dist.new_empty((dist.shape[0],), dtype=torch.uint8))


@_op("representatives_gram")
def representatives_gram(G: T_, event_offsets: T_, order: T_, k: int, n_rep: int) -> Tuple[T_, T_, T_, T_]:
    return ops.representatives_gram(G, event_offsets, order, k, n_rep)


@representatives_gram.register_fake
def _(G, event_offsets, order, k, n_rep):
    T = G.shape[0]
    return (G.new_empty((T,), dtype=torch.float32), G.new_empty((k, n_rep), dtype=torch.int64), G.new_empty((k,), dtype=torch.int32),
            G.new_empty((T,), dtype=torch.uint8))


@_op("event_order")
def event_order(assignments: T_, k: int) -> Tuple[T_, T_]:
    return ops.event_order(assignments, k)


@event_order.register_fake
def _(assignments, k):
    return assignments.new_empty(assignments.shape, dtype=torch.int32), assignments.new_empty((k + 1,), dtype=torch.int32)


@_op("gather_rows")
def gather_rows(src: T_, index: T_) -> T_:
    out = torch.empty(index.shape[0], src.shape[1], dtype=src.dtype, device=src.device)
    return ops.gather_rows(src, index, out, n=index.shape[0])


@gather_rows.register_fake
def _(src, index):
    return src.new_empty((index.shape[0], src.shape[1]))


# ------------------------------------------------------------------------------------------------ (c) relevance
@_op("cosine_flags")
def cosine_flags(question: T_, events: T_, threshold: float = 0.45) -> Tuple[T_, T_]:
    return ops.cosine_flags(question, events, threshold)


@cosine_flags.register_fake
def _(question, events, threshold=0.45):
    return events.new_empty((events.shape[0],)), events.new_empty((events.shape[0],), dtype=torch.uint8)


@_op("minor_frames")
def minor_frames(assignments: T_, low_flags: T_, rep_flags: T_) -> Tuple[T_, T_, T_]:
    return ops.minor_frames(assignments, low_flags, rep_flags)


@minor_frames.register_fake
def _(assignments, low_flags, rep_flags):
    T = assignments.shape[0]
    return (assignments.new_empty((T,), dtype=torch.uint8), assignments.new_empty((T,), dtype=torch.int32),
            assignments.new_empty((1,), dtype=torch.int64))


@_op("relevance_minor", mutates=("ticket",))
def relevance_minor(question: T_, events: T_, assignments: T_, rep_flags: T_, ticket: T_,
                    threshold: float = 0.45) -> Tuple[T_, T_, T_, T_, T_]:
    return ops.relevance_minor(question, events, assignments, rep_flags, ticket, threshold)


@relevance_minor.register_fake
def _(question, events, assignments, rep_flags, ticket, threshold=0.45):
    k, T = events.shape[0], assignments.shape[0]
    return (events.new_empty((k,)), events.new_empty((k,), dtype=torch.uint8), events.new_empty((T,), dtype=torch.uint8),
            events.new_empty((T,), dtype=torch.int32), events.new_empty((1,), dtype=torch.int64))


_SEL_KEYS = ("cos", "low", "dist", "reps", "rep_counts", "rep_flags", "minor_flags", "minor_list", "n_minor")


def _sel_fake(src, k, T, n_rep):
    return (src.new_empty((k,)), src.new_empty((k,), dtype=torch.uint8), src.new_empty((T,), dtype=torch.float32),
            src.new_empty((k, n_rep), dtype=torch.int64), src.new_empty((k,), dtype=torch.int32), src.new_empty((T,), dtype=torch.uint8),
            src.new_empty((T,), dtype=torch.uint8), src.new_empty((T,), dtype=torch.int32), src.new_empty((1,), dtype=torch.int64))


@_op("relevance_select", mutates=("ticket",))
def relevance_select(question: T_, events: T_, G: T_, event_offsets: T_, event_order_: T_, assignments: T_, ticket: T_, n_rep: int = 2,
                     threshold: float = 0.45) -> Tuple[T_, T_, T_, T_, T_, T_, T_, T_, T_]:
    """-> (cos, low, dist, reps, rep_counts, rep_flags, minor_flags, minor_list, n_minor)."""
    b = ops.relevance_select(question, events, G, event_offsets, event_order_, assignments, ticket, k=events.shape[0], n_rep=n_rep,
                             threshold=threshold)
    return tuple(b[n] for n in _SEL_KEYS)


@relevance_select.register_fake
def _(question, events, G, event_offsets, event_order_, assignments, ticket, n_rep=2, threshold=0.45):
    return _sel_fake(events, events.shape[0], assignments.shape[0], n_rep)


@_op("relevance_select_hidden", mutates=("ticket",))
def relevance_select_hidden(question: T_, hidden: T_, cu_seqlens: T_, G: T_, event_offsets: T_, event_order_: T_, assignments: T_,
                            ticket: T_, n_rep: int = 2, threshold: float = 0.45) -> Tuple[T_, T_, T_, T_, T_, T_, T_, T_, T_, T_]:
    """Event embeddings = mean pool of the packed hidden states, fused in.  -> (..., events [k, Dl])."""
    k = cu_seqlens.shape[0] - 1
    b = ops.relevance_select(question, None, G, event_offsets, event_order_, assignments, ticket, k=k, n_rep=n_rep, threshold=threshold,
                             hidden=hidden, cu_seqlens=cu_seqlens)
    return tuple(b[n] for n in _SEL_KEYS) + (b["events"],)


@relevance_select_hidden.register_fake
def _(question, hidden, cu_seqlens, G, event_offsets, event_order_, assignments, ticket, n_rep=2, threshold=0.45):
    k = cu_seqlens.shape[0] - 1
    return _sel_fake(hidden, k, assignments.shape[0], n_rep) + (hidden.new_empty((k, hidden.shape[1])),)


# ------------------------------------------------------------------------------------------------ (a), (d) masks / pooling / compaction
@_op("pool_norm")
def pool_norm(x: T_) -> Tuple[T_, T_]:
    pb, _, sq = ops.pool_norm(x)
    return pb, sq


@pool_norm.register_fake
def _(x):
    return x.new_empty((x.shape[0], x.shape[2])), x.new_empty((x.shape[0],), dtype=torch.float32)


@_op("pool_frames_inplace", mutates=("x",))
def pool_frames_inplace(x: T_, P: int, frames: T_) -> None:
    ops.pool_frames_inplace(x, P, frames)


@pool_frames_inplace.register_fake
def _(x, P, frames):
    return None


@_op("pixel_diff")
def pixel_diff(pixels: T_, t: int, P: int, W: int, threshold: float = 0.1, max_ctas: int = 0, evict_first: bool = True) -> T_:
    return ops.pixel_diff(pixels, t, P, W, threshold=threshold, max_ctas=max_ctas, evict_first=evict_first)


@pixel_diff.register_fake
def _(pixels, t, P, W, threshold=0.1, max_ctas=0, evict_first=True):
    return pixels.new_empty((t * P,), dtype=torch.uint8)


@_op("mask_finalize", mutates=("mask",))
def mask_finalize(mask: T_, t: int, P: int, min_tokens: int = 1, minor_flags: Optional[T_] = None, frame0: int = 0) -> None:
    ops.mask_finalize(mask, t, P, min_tokens=min_tokens, minor_flags=minor_flags, frame0=frame0)


@mask_finalize.register_fake
def _(mask, t, P, min_tokens=1, minor_flags=None, frame0=0):
    return None


@_op("pixel_mask")
def pixel_mask(pixels: T_, t: int, P: int, W: int, threshold: float = 0.1, min_tokens: int = 1,
               minor_flags: Optional[T_] = None, frame0: int = 0) -> T_:
    return ops.pixel_mask(pixels, t, P, W, threshold=threshold, min_tokens=min_tokens, minor_flags=minor_flags, frame0=frame0)


@pixel_mask.register_fake
def _(pixels, t, P, W, threshold=0.1, min_tokens=1, minor_flags=None, frame0=0):
    return pixels.new_empty((t * P,), dtype=torch.uint8)


@_op("compact_tokens")
def compact_tokens(x: T_, T: int, P: int, mask: T_, pool_flags: Optional[T_] = None) -> Tuple[T_, T_, T_]:
    return ops.compact_tokens(x, T, P, mask, pool_flags)


@compact_tokens.register_fake
def _(x, T, P, mask, pool_flags=None):
    n = T * P
    return x.new_empty((n, x.shape[1])), x.new_empty((n,), dtype=torch.int32), x.new_empty((1,), dtype=torch.int64)


@_op("filter_ids")
def filter_ids(input_ids: T_, image_token_index: int, mask: T_, attention_mask: Optional[T_] = None) -> Tuple[T_, T_, T_]:
    """-> (out_ids, out_attention_mask (zeros when none given), n_out [2] = {ids kept, <image> ids seen})."""
    o_ids, o_am, _, _, n_out = ops.filter_ids(input_ids, image_token_index, mask, attention_mask)
    return o_ids, o_am if o_am is not None else torch.zeros_like(o_ids), n_out


@filter_ids.register_fake
def _(input_ids, image_token_index, mask, attention_mask=None):
    return torch.empty_like(input_ids), torch.empty_like(input_ids), input_ids.new_empty((2,))


@_op("compact_stream", mutates=("mask",))
def compact_stream(x: T_, T: int, P: int, mask: T_, minor_flags: Optional[T_], input_ids: T_, image_token_index: int,
                   attention_mask: Optional[T_] = None, min_tokens: int = 1, finalize: bool = True) -> Tuple[T_, T_, T_, T_, T_, T_]:
    """Per-frame mask rules (in place on ``mask``) + compaction with pooled minor frames + id filter.
    -> (out [T*P, D], kept_index, n_kept [1], out_ids, out_attention (zeros when none given), n_out [2])."""
    b = ops.compact_stream(x, T, P, mask, minor_flags, min_tokens=min_tokens, finalize=finalize, input_ids=input_ids,
                           image_token_index=image_token_index, attention_mask=attention_mask)
    am = b.get("out_attention")
    return b["out"], b["kept_index"], b["n_kept"], b["out_ids"], am if am is not None else torch.zeros_like(b["out_ids"]), b["n_out"]


@compact_stream.register_fake
def _(x, T, P, mask, minor_flags, input_ids, image_token_index, attention_mask=None, min_tokens=1, finalize=True):
    n = T * P
    return (x.new_empty((n, x.shape[1])), x.new_empty((n,), dtype=torch.int32), x.new_empty((1,), dtype=torch.int64),
            torch.empty_like(input_ids), torch.empty_like(input_ids), input_ids.new_empty((2,)))


# ------------------------------------------------------------------------------------------------ around the frozen LLM
@_op("embed_scatter")
def embed_scatter(input_ids: T_, image_token_index: int, embed_weight: T_, features: T_) -> Tuple[T_, T_]:
    return ops.embed_scatter(input_ids, image_token_index, embed_weight, features, check_count=False)


@embed_scatter.register_fake
def _(input_ids, image_token_index, embed_weight, features):
    return embed_weight.new_empty((input_ids.shape[0], embed_weight.shape[1])), input_ids.new_empty((2,), dtype=torch.int32)


@_op("embed_scatter_ordered")
def embed_scatter_ordered(input_ids: T_, image_token_index: int, embed_weight: T_, features: T_, frame_order: T_,
                          tokens_per_frame: int) -> Tuple[T_, T_]:
    return ops.embed_scatter(input_ids, image_token_index, embed_weight, features, check_count=False, frame_order=frame_order,
                             tokens_per_frame=tokens_per_frame)


@embed_scatter_ordered.register_fake
def _(input_ids, image_token_index, embed_weight, features, frame_order, tokens_per_frame):
    return embed_weight.new_empty((input_ids.shape[0], embed_weight.shape[1])), input_ids.new_empty((2,), dtype=torch.int32)


@_op("seq_mean_pool")
def seq_mean_pool(hidden: T_, cu_seqlens: T_) -> T_:
    return ops.seq_mean_pool(hidden, cu_seqlens)


@seq_mean_pool.register_fake
def _(hidden, cu_seqlens):
    return hidden.new_empty((cu_seqlens.shape[0] - 1, hidden.shape[1]))


@_op("ln_merge")
def ln_merge(hidden: T_, merge: int, gamma: T_, beta: T_, eps: float) -> T_:
    return ops.ln_merge(hidden, merge, gamma, beta, eps)


@ln_merge.register_fake
def _(hidden, merge, gamma, beta, eps):
    return hidden.new_empty((hidden.shape[0] // (merge * merge), hidden.shape[1]))
