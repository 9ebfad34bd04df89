"""Per-stream device cache for incremental re-clustering.

The reference re-decodes, re-encodes and re-clusters the whole history for every new question
(``evaluate/answer_generate.py:120-135``, ``model/processing_cogreasoner.py:468-507``).  With the Gram formulation
the expensive part of a re-clustering is G = X X^T, and G only grows: appending frames [T_old, T) needs the tiles that
touch the new frames -- work ~ (T - T_old) * T * K instead of T^2 * K -- after which the k-means (new k = ceil(T/15), fresh
init, as the reference does) runs on the cached matrix.  Results are those of a from-scratch run on the same T frames:
same algorithm, same Gram entries up to the fp32 partial-sum grouping.
"""
from __future__ import annotations

import math
import random
from typing import Optional, Sequence

import torch

from . import ops


class StreamCache:
    """Append-only feature / timestamp / Gram cache of ONE stream (``capacity`` frames of P tokens x D channels)."""

    def __init__(self, capacity: int, P: int, D: int, *, device, ts_dtype=torch.float64):
        self.capacity, self.P, self.D, self.K = capacity, P, D, P * D
        self.device = torch.device(device)
        self.X = torch.empty(capacity, self.K, dtype=torch.bfloat16, device=self.device)
        self.ts = torch.empty(capacity, dtype=ts_dtype, device=self.device)
        self.G = torch.empty(capacity, capacity, dtype=torch.float64, device=self.device)
        self.ws = torch.empty(ops.gram_workspace_bytes(capacity, self.K), dtype=torch.uint8, device=self.device)
        self.rng_state = torch.randint(0, 2 ** 62, (2,), dtype=torch.int64, device=self.device)  # the solver's built-in generator
        self.rng_state[1] = 0
        self.T = 0

    def append(self, features: torch.Tensor, timestamps: torch.Tensor, *, gram_impl: int = 0) -> int:
        """features [t*P, D] (or [t, P, D]) bf16 of the NEW frames, timestamps [t]; returns the new frame count."""
        f = features.reshape(-1, self.K)
        t = f.shape[0]
        if self.T + t > self.capacity:
            raise ValueError(f"StreamCache: {self.T} + {t} frames exceed the capacity {self.capacity}")
        self.X[self.T: self.T + t].copy_(f)
        self.ts[self.T: self.T + t].copy_(timestamps)
        T_new = self.T + t
        ops.gram_extend(self.X[:T_new], self.T, self.G, impl=gram_impl, workspace=self.ws)
        self.T = T_new
        return self.T

    @property
    def gram(self) -> torch.Tensor:
        return self.G[: self.T, : self.T]

    def cluster(self, cluster_num: Optional[int] = None, *, alpha=2, max_iteration=30, tol=1e-4,
                init_indices: Optional[Sequence[int]] = None, first_index: Optional[int] = None,
                exp_noise: Optional[torch.Tensor] = None, reseed_pool: Optional[torch.Tensor] = None, validate: bool = False):
        """k-means over all cached frames on the cached Gram matrix (k = ceil(T/15) by default, cogreasoner_chat.py:280).
        Returns ``(KMeansBuffers, k)``: assignments, centre CSR, event order etc. stay on the device (no host sync)."""
        T = self.T
        k = int(cluster_num) if cluster_num is not None else math.ceil(T / 15)
        if T <= k:
            raise ValueError("StreamCache.cluster: need more frames than clusters (the reference returns early, kmeans_with_time.py:30)")
        buf = ops.KMeansBuffers(T, k, self.ts.dtype, self.device)
        noise, rng = None, None
        if init_indices is not None:
            buf.init_indices.copy_(torch.as_tensor(list(init_indices), dtype=torch.int64))
            n_given = k
        elif first_index is None and exp_noise is None:
            n_given, rng = 0, self.rng_state  # first centre, Exp(1) noise and re-seeds from the solver's built-in generator
        else:
            first = random.randint(0, T - 1) if first_index is None else int(first_index)
            buf.init_indices[:1].fill_(first)
            n_given = 1
            if k > 1:
                noise = exp_noise if exp_noise is not None else torch.empty(k - 1, T, dtype=torch.float32, device=self.device).exponential_(1)
        # re-seed draws: the built-in generator cannot run out; a caller-supplied pool can (`validate=True` reads the status back,
        # one sync, and raises on overflow); without either, a pool of the strict upper bound k * (max_iteration + 1) is drawn
        pool = reseed_pool
        if pool is None and rng is None:
            if init_indices is not None:
                rng = self.rng_state
            else:
                pool = torch.randint(0, T, (k * (int(max_iteration) + 1),), dtype=torch.int64, device=self.device)
        ops.kmeans_solve(self.gram, self.ts[:T], buf, n_init_given=n_given, exp_noise=noise, reseed_pool=pool, alpha=float(alpha),
                         max_iteration=int(max_iteration), tol=float(tol), rng_state=rng)
        if validate and ops.decode_status(buf.status)["reseed_overflow"]:
            raise RuntimeError("StreamCache.cluster: reseed_pool exhausted; the clustering is not the reference's")
        return buf, k

    def centroids(self, buf: ops.KMeansBuffers, k: int) -> torch.Tensor:
        return ops.centroid_update(self.X[: self.T], buf.member_offsets, buf.members, k).view(k, self.P, self.D)
