"""Drop-in for the reference's ``model/kmeans_with_time.py`` (same name, same signature).

    from cogstream_b200.kmeans_with_time import kmeans_with_time_min_max

Reference: ``kmeans_with_time_min_max(features, timestamp, cluster_num, alpha=2, max_iteration=30,
tol=1e-4)`` (model/kmeans_with_time.py:4).  Same return contract: ``(features_kmeans [k,P,D] fp32,
times_kmeans [k] (timestamp dtype), cluster_assignments [T] int64)`` and the ``T <= k`` early
return ``(features_fp32, timestamp[:k], None)`` (:30-32).

How it runs on B200: ONE tcgen05 pass builds the T x T Gram matrix of the bf16 frame vectors; the
k-means++ init and every Lloyd iteration are then evaluated on that matrix by a single cooperative
kernel (``csrc/kmeans_solver.cu``); one last streaming pass produces the fp32 centroids.

Host RNG parity: the reference consumes ``random.randint`` for the first centre (:41),
for degenerate init rounds (:57) and for empty clusters (:118), and ``torch.multinomial`` on the
features' device for the other centres (:60).  This wrapper makes the same draws from the same
generators in the same order -- ``multinomial(probs, 1)`` is ``argmax(probs / q)``, ``q ~ Exp(1)``
drawn by ``empty_like(probs).exponential_(1)``, so the noise is pre-drawn with exactly those calls and
the generators are rewound to what the reference would have consumed.  Keyword-only ``init_indices``,
``exp_noise`` and ``reseed_indices`` inject the draws instead (parity tests).
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import ops

_RESEED_POOL = 64


@dataclass
class KMeansRunInfo:
    init_indices: torch.Tensor  # [k] int64 (device)
    iterations: int
    converged: bool
    n_reseed_used: int
    degenerate_rounds: int
    n_noise_used: int
    min_margin: float
    last_shift: float
    member_offsets: torch.Tensor  # [k+1] int32: CSR of the frames defining each returned centre
    members: torch.Tensor  # [T+k] int32
    gram: torch.Tensor  # [T,T] fp64


def _as_bf16_frames(features: torch.Tensor) -> torch.Tensor:
    if not isinstance(features, torch.Tensor):
        raise TypeError("features must be a torch.Tensor on a CUDA device")
    if not features.is_cuda:
        raise RuntimeError("kmeans_with_time_min_max: CPU tensor given; cogstream_b200 has no CPU fallback")
    if features.dim() != 3:
        raise ValueError(f"features must be [T, P, D], got {tuple(features.shape)}")
    if not features.is_contiguous():
        # the reference's .view(T, P*D) (kmeans_with_time.py:35) has the same requirement
        raise RuntimeError("features must be contiguous (the reference's view(T, P*D) requires it too)")
    if features.dtype == torch.bfloat16:
        return features
    if features.dtype in (torch.float32, torch.float16, torch.float64):
        as_bf16 = features.to(torch.bfloat16)
        if not torch.equal(as_bf16.to(features.dtype), features):
            raise NotImplementedError(
                "kmeans_with_time_min_max: features are not exactly representable in bf16; the B200 path consumes the "
                "bf16 encoder/projector output (cogreasoner_chat.py:264-276) and has no wider-precision kernel")
        return as_bf16
    raise TypeError(f"unsupported feature dtype {features.dtype}")


def kmeans_with_time_min_max(features, timestamp, cluster_num, alpha=2, max_iteration=30, tol=1e-4, *,
                             init_indices: Optional[Sequence[int]] = None, exp_noise: Optional[torch.Tensor] = None,
                             reseed_indices: Optional[Sequence[int]] = None, first_index: Optional[int] = None,
                             gram_impl: int = 0, return_info: bool = False, return_centroids: bool = True):
    """See module docstring.  Extra keyword-only arguments default to the reference behaviour.  ``return_centroids=False``
    skips the streaming pass that produces the fp32 centroids (first element ``None``): callers that select from the Gram matrix
    (``select_events_based_on_summary``) never read them."""
    x3 = _as_bf16_frames(features)
    dev = x3.device
    if not isinstance(timestamp, torch.Tensor):  # kmeans_with_time.py:26-27
        timestamp = torch.tensor(timestamp, dtype=torch.float32)
    T, P, D = x3.shape
    k = int(cluster_num)
    if T <= k:  # :30-32
        out = (features.to(dtype=torch.float32), timestamp[:k], None)
        return out + (None,) if return_info else out
    if timestamp.dtype not in (torch.float32, torch.float64):
        timestamp = timestamp.to(torch.float32 if not timestamp.dtype.is_floating_point else torch.float64)
    ts = timestamp.to(dev).contiguous()
    if ts.shape[0] != T:
        raise ValueError(f"timestamp has {ts.shape[0]} entries for {T} frames")

    x2 = x3.view(T, P * D)
    G = ops.gram(x2, impl=gram_impl)
    buf = ops.KMeansBuffers(T, k, ts.dtype, dev)

    # ---- host RNG draws, in the reference's order ----------------------------------------------
    py_state = None
    gen_state = None
    noise = None
    if init_indices is not None:
        idx = torch.as_tensor(list(init_indices), dtype=torch.int64)
        if idx.numel() != k:
            raise ValueError("init_indices must have cluster_num entries")
        buf.init_indices.copy_(idx)
        n_given = k
    else:
        first = random.randint(0, T - 1) if first_index is None else int(first_index)  # :41
        buf.init_indices[0] = first
        n_given = 1
        if k > 1:
            if exp_noise is not None:
                noise = exp_noise.to(device=dev, dtype=torch.float32).contiguous()
            else:
                gen_state = torch.cuda.get_rng_state(dev)
                noise = torch.empty(k - 1, T, dtype=torch.float32, device=dev)
                for m in range(k - 1):  # one [T] draw per multinomial call, as ATen does (:60)
                    noise[m].exponential_(1)
    if reseed_indices is not None:
        pool_host = [int(v) for v in reseed_indices]
    else:
        py_state = random.getstate()
        pool_host = [random.randint(0, T - 1) for _ in range(_RESEED_POOL)]  # speculative; rewound below

    while True:
        pool = torch.tensor(pool_host, dtype=torch.int64, device=dev) if pool_host else None
        ops.kmeans_solve(G, ts, buf, n_init_given=n_given, exp_noise=noise, reseed_pool=pool, alpha=float(alpha),
                         max_iteration=int(max_iteration), tol=float(tol))
        st = ops.decode_status(buf.status)  # device -> host sync (the reference syncs every iteration)
        if not st["reseed_overflow"]:
            break
        if reseed_indices is not None:
            raise RuntimeError("kmeans_with_time_min_max: reseed_indices exhausted")
        random.setstate(py_state)
        pool_host = [random.randint(0, T - 1) for _ in range(len(pool_host) * 4)]
        if init_indices is None:
            buf.init_indices[0] = first

    # ---- rewind the generators to exactly what the reference would have consumed ------------------
    if py_state is not None:
        random.setstate(py_state)
        for _ in range(st["n_reseed_used"]):
            random.randint(0, T - 1)
    if gen_state is not None and st["n_noise_used"] != k - 1:
        torch.cuda.set_rng_state(gen_state, dev)
        scratch = torch.empty(T, dtype=torch.float32, device=dev)
        for _ in range(st["n_noise_used"]):
            scratch.exponential_(1)

    centroids = ops.centroid_update(x2, buf.member_offsets, buf.members, k).view(k, P, D) if return_centroids else None  # :134
    out = (centroids, buf.centroid_times, buf.assignments)
    if return_info:
        info = KMeansRunInfo(buf.init_indices, st["iterations"], st["converged"], st["n_reseed_used"], st["degenerate_rounds"],
                             st["n_noise_used"], st["min_margin"], st["last_shift"], buf.member_offsets, buf.members, G)
        return out + (info,)
    return out
