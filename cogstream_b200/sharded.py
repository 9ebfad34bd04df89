"""Frame-sharded k-means for ONE long stream over the GPUs of a box (first half; BASELINE config 4).

The reference clusters a stream on one device (``model/kmeans_with_time.py``); a 2-hour stream (T = 14,400 frames, k = 960) keeps
seven of eight GPUs idle.  Here every rank holds the stream's Gram matrix (``StreamCache`` keeps it current: appending a segment
costs the new tiles only) and

  * runs the k-means++ init redundantly -- ``k - 1`` dependent sampling rounds, a latency chain with nothing to shard; all ranks use
    the same generator state, so they pick the same centres;
  * splits the Lloyd iteration by FRAMES: rank ``r`` owns frames ``[r * per, (r + 1) * per)`` and computes their distance /
    assignment pass and their rows of the sum-of-Gram-columns matrix S only.  Two NCCL all-gathers per iteration (over NVLink):
    the int32 assignments ``[T]`` and, per frame, the two S entries the centre-shift needs ``[world * (2 per + k + 1)]`` fp64;
    membership, centre times, re-seeds, ||c||^2 and the convergence test are evaluated identically on every rank.

Same device functions and summation orders as the single-GPU solver (``csb_kmeans_solve``): assignments, iteration count,
re-seed count and centre times are bit-identical to it (``tests/test_gpu_parity.py::test_frame_sharded_lloyd_*``).

    km = FrameShardedKMeans(T, k, torch.float64, device, rank=rank, world=world)
    buf = km.solve(cache.gram, cache.ts[:T], rng_state=shared_state)      # collective: every rank calls it
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Iterator, Optional, Tuple

import torch

from . import _lib, ops
from ._lib import KMeansParams, LloydShard, check

__all__ = ["FrameShardedKMeans", "frame_range"]


def frame_range(T: int, rank: int, world: int) -> Tuple[int, int, int]:
    """-> (per_rank, f_begin, f_end): contiguous frame slices of ``ceil(T / world)`` frames (the last ranks may own fewer / none)."""
    per = math.ceil(T / world)
    b = min(T, rank * per)
    return per, b, min(T, b + per)


class FrameShardedKMeans:
    """Buffers + phase driver of one rank.  ``solve`` is a collective call (every rank, same arguments)."""

    def __init__(self, T: int, k: int, ts_dtype: torch.dtype, device, *, rank: int = 0, world: int = 1, group=None):
        if T <= k:
            raise ValueError("FrameShardedKMeans: need more frames than clusters (the reference returns early, kmeans_with_time.py:30)")
        self.T, self.k, self.rank, self.world, self.group = T, k, rank, world, group
        self.device = torch.device(device)
        self.per, self.f_begin, self.f_end = frame_range(T, rank, world)
        self.buf = ops.KMeansBuffers(T, k, ts_dtype, self.device)
        self.assign32 = torch.zeros(world * self.per, dtype=torch.int32, device=self.device)
        self.seg = 2 * self.per + k + 1
        self.exchange = torch.zeros(world * self.seg, dtype=torch.float64, device=self.device)
        self.state = torch.zeros(int(_lib.load().csb_kmeans_lloyd_state_bytes(T, k)), dtype=torch.uint8, device=self.device)
        self._flags = self.state[-256:].view(torch.int32)[:8]
        self._flags_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self.collective_ms = 0.0
        self.iterations = 0

    # ----------------------------------------------------------------------------------------------------------------
    def _phase(self, phase: int, G, ts, p: KMeansParams, pool: Optional[torch.Tensor]):
        b = self.buf
        sh = LloydShard(phase, 0, self.f_begin, self.f_end, self.per, self.exchange.data_ptr(), self.state.data_ptr(), self.state.numel())
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(_lib.load().csb_kmeans_lloyd_phase(
            G.data_ptr(), ts.data_ptr(), C.byref(p), C.byref(sh), b.init_indices.data_ptr(), None if pool is None else pool.data_ptr(),
            self.assign32.data_ptr(), b.assignments.data_ptr(), b.centroid_times.data_ptr(), b.member_offsets.data_ptr(),
            b.members.data_ptr(), b.event_order.data_ptr(), b.event_offsets.data_ptr(), b.status.data_ptr(), b.workspace.data_ptr(),
            b.workspace.numel(), stream), "kmeans_lloyd_phase")

    def phases(self, G: torch.Tensor, timestamps: torch.Tensor, *, init_indices=None, first_index: Optional[int] = None,
               exp_noise: Optional[torch.Tensor] = None, reseed_pool: Optional[torch.Tensor] = None,
               rng_state: Optional[torch.Tensor] = None, alpha: float = 2.0, max_iteration: int = 30,
               tol: float = 1e-4) -> Iterator[Tuple[torch.Tensor, int]]:
        """The phase sequence as a generator: yields ``(buffer, count_per_rank)`` whenever an all-gather of ``buffer`` (rank-major,
        ``count_per_rank`` elements per rank, this rank's slice filled) has to happen before it is resumed.  ``solve`` drives it with
        ``torch.distributed``; the tests drive several emulated ranks in lock step on one GPU."""
        T, k = self.T, self.k
        _, ldg = ops._gram_view(G)
        b = self.buf
        n_pool = 0 if reseed_pool is None else reseed_pool.numel()
        if init_indices is not None:
            b.init_indices.copy_(torch.as_tensor(init_indices, dtype=torch.int64))
            n_given = k
        elif first_index is not None:
            b.init_indices[:1].fill_(int(first_index))
            n_given = 1
        else:
            n_given = 0
        if n_given < k and exp_noise is None and rng_state is None:
            raise ValueError("FrameShardedKMeans: sampled centres need exp_noise or a (shared) rng_state")
        if reseed_pool is None and rng_state is None:
            raise ValueError("FrameShardedKMeans: re-seeds need reseed_pool or a (shared) rng_state")
        # k-means++ (replicated): init only
        ops.kmeans_solve(G, timestamps, b, n_init_given=n_given, exp_noise=exp_noise, reseed_pool=reseed_pool, alpha=alpha,
                         max_iteration=-1, tol=tol, rng_state=rng_state)
        p = KMeansParams(T, k, ops._TS_DTYPE[timestamps.dtype], int(max_iteration), float(alpha), float(tol), n_given, n_pool, ldg,
                         None if rng_state is None else rng_state.data_ptr(), 0)
        self._phase(_lib.LLOYD_INIT, G, timestamps, p, reseed_pool)
        it, converged = 0, False
        while it < max_iteration and not converged:
            self._phase(_lib.LLOYD_ASSIGN, G, timestamps, p, reseed_pool)
            yield self.assign32, self.per
            self._phase(_lib.LLOYD_UPDATE, G, timestamps, p, reseed_pool)
            yield self.exchange, self.seg
            self._phase(_lib.LLOYD_SHIFT, G, timestamps, p, reseed_pool)
            self._flags_host.copy_(self._flags, non_blocking=True)  # the reference syncs on `shift <= tol` every iteration too (:130)
            torch.cuda.current_stream(self.device).synchronize()
            it, converged = int(self._flags_host[1]), bool(self._flags_host[3])
        self.iterations = it
        self._phase(_lib.LLOYD_OUTPUT, G, timestamps, p, reseed_pool)

    def solve(self, G: torch.Tensor, timestamps: torch.Tensor, **kw) -> ops.KMeansBuffers:
        """Collective: every rank calls it with the same Gram matrix / timestamps / randomness.  Returns this rank's
        ``KMeansBuffers`` (assignments, centre times, CSR, event order, status: complete and identical on every rank)."""
        import torch.distributed as dist
        self.collective_ms = 0.0
        timed = []
        for buffer, count in self.phases(G, timestamps, **kw):
            if self.world > 1:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dist.all_gather_into_tensor(buffer, buffer[self.rank * count: (self.rank + 1) * count], group=self.group)
                e1.record()
                timed.append((e0, e1))
        if timed:
            torch.cuda.current_stream(self.device).synchronize()
            self.collective_ms = sum(a.elapsed_time(b) for a, b in timed)
        return self.buf
