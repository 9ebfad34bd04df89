"""torch.library registration of the C-ABI kernels (namespace ``cogstream_b200``).

The ops are thin: each implementation forwards tensors' ``data_ptr()`` / sizes / the current stream to the shared
library through :mod:`cogstream_b200.ops`.  CUDA only (``device_types="cuda"``): calling one with CPU tensors raises
``NotImplementedError`` from the dispatcher -- there is no CPU kernel to fall back to.  Fake (meta) implementations
give shapes/dtypes so the ops compose with ``torch.compile`` / export without running anything.

    torch.ops.cogstream_b200.gram(x)                         -> G [T,T] f64
    torch.ops.cogstream_b200.centroid_update(x, offs, members, k) -> [k,K] f32
    torch.ops.cogstream_b200.dist_to_assigned(x, centroids, assign) -> [T] f32
    torch.ops.cogstream_b200.pool_norm(x)                    -> (pooled bf16 [T,D], sqnorm f32 [T])
    torch.ops.cogstream_b200.cosine_flags(q, events, thr)    -> (cos [k], low u8 [k])
    torch.ops.cogstream_b200.pixel_mask(pix, t, P, W, thr, min_tokens, minor_flags?, frame0) -> u8 [t*P]
    torch.ops.cogstream_b200.compact_tokens(x, T, P, mask, pool_flags?) -> (out [T*P,D], kept_index i32 [T*P], n_kept i64 [1])
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.library import custom_op

from . import ops

NS = "cogstream_b200"


@custom_op(f"{NS}::gram", mutates_args=(), device_types="cuda")
def gram(x: torch.Tensor, impl: int = 0) -> torch.Tensor:
    return ops.gram(x, impl=impl)


@gram.register_fake
def _(x, impl=0):
    return x.new_empty((x.shape[0], x.shape[0]), dtype=torch.float64)


@custom_op(f"{NS}::centroid_update", mutates_args=(), device_types="cuda")
def centroid_update(x: torch.Tensor, member_offsets: torch.Tensor, members: torch.Tensor, k: int) -> torch.Tensor:
    return ops.centroid_update(x, member_offsets, members, k)


@centroid_update.register_fake
def _(x, member_offsets, members, k):
    return x.new_empty((k, x.shape[1]), dtype=torch.float32)


@custom_op(f"{NS}::dist_to_assigned", mutates_args=(), device_types="cuda")
def dist_to_assigned(x: torch.Tensor, centroids: torch.Tensor, assignments: torch.Tensor) -> torch.Tensor:
    return ops.dist_to_assigned(x, centroids, assignments)


@dist_to_assigned.register_fake
def _(x, centroids, assignments):
    return x.new_empty((x.shape[0],), dtype=torch.float32)


@custom_op(f"{NS}::pool_norm", mutates_args=(), device_types="cuda")
def pool_norm(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    pb, _, sq = ops.pool_norm(x)
    return pb, sq


@pool_norm.register_fake
def _(x):
    return x.new_empty((x.shape[0], x.shape[2])), x.new_empty((x.shape[0],), dtype=torch.float32)


@custom_op(f"{NS}::cosine_flags", mutates_args=(), device_types="cuda")
def cosine_flags(question: torch.Tensor, events: torch.Tensor, threshold: float = 0.45) -> Tuple[torch.Tensor, torch.Tensor]:
    return ops.cosine_flags(question, events, threshold)


@cosine_flags.register_fake
def _(question, events, threshold=0.45):
    return events.new_empty((events.shape[0],)), events.new_empty((events.shape[0],), dtype=torch.uint8)


@custom_op(f"{NS}::pixel_mask", mutates_args=(), device_types="cuda")
def pixel_mask(pixels: torch.Tensor, t: int, P: int, W: int, threshold: float = 0.1, min_tokens: int = 1,
               minor_flags: Optional[torch.Tensor] = None, frame0: int = 0) -> torch.Tensor:
    return ops.pixel_mask(pixels, t, P, W, threshold=threshold, min_tokens=min_tokens, minor_flags=minor_flags, frame0=frame0)


@pixel_mask.register_fake
def _(pixels, t, P, W, threshold=0.1, min_tokens=1, minor_flags=None, frame0=0):
    return pixels.new_empty((t * P,), dtype=torch.uint8)


@custom_op(f"{NS}::compact_tokens", mutates_args=(), device_types="cuda")
def compact_tokens(x: torch.Tensor, T: int, P: int, mask: torch.Tensor,
                   pool_flags: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    return ops.compact_tokens(x, T, P, mask, pool_flags)


@compact_tokens.register_fake
def _(x, T, P, mask, pool_flags=None):
    n = T * P
    return x.new_empty((n, x.shape[1])), x.new_empty((n,), dtype=torch.int32), x.new_empty((1,), dtype=torch.int64)
