"""Fused Visual Stream Compression pipeline: the same path as the drop-in methods, but enqueued
back-to-back on one CUDA stream with pre-allocated buffers and no host round trip inside a phase.

    phase 1  (before the frozen LLM)  cogreasoner_chat.py:279-301
        Gram (tcgen05) -> k-means++ + Lloyd (one cooperative kernel) -> fp32 centroids ->
        chronological event order -> representatives -> packed per-event token gather
    -- frozen LLM forwards (:302-323), untouched: event / question embeddings come back --
    phase 2                          cogreasoner_chat.py:325-333, :551-564
        cosine flags -> minor frames -> pixel-difference mask (+ min_tokens, minor override) ->
        scan + compaction (minor frames emitted as their pooled token) -> input-id filtering

Independent streams shard across ranks (``shard_streams``); the only collective is one all-gather of
the per-stream stats rows (``gather_stats``), as in SURVEY.md section 8(e).
"""
from __future__ import annotations

import math
import os
import random
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import ops

STATS_FIELDS = ("T", "k", "iterations", "n_reseed", "n_minor", "n_kept", "n_ids_out", "converged", "reseed_overflow")


@dataclass
class StreamConfig:
    """Hot-path constants, reference values as defaults."""

    frames_per_event: int = 15  # cogreasoner_chat.py:280
    min_events: int = 9  # :281
    n_representatives: int = 2  # :284
    relevance_threshold: float = 0.45  # :329
    pixel_threshold: float = 0.1  # :390
    min_tokens: int = 1  # :391
    alpha: float = 2.0  # kmeans_with_time.py:4
    max_iteration: int = 30
    tol: float = 1e-4
    cluster_num: Optional[int] = None  # override k = ceil(T / frames_per_event) and its gate
    gram_impl: int = 0
    # entries of the device-drawn re-seed pool (random.randint draws of kmeans_with_time.py:57 / :118).  None = the strict upper
    # bound k * (max_iteration + 1): at most k - 1 degenerate init rounds plus k empty clusters per iteration, so the pool can
    # never run out.  A caller-supplied smaller pool that does run out sets `reseed_overflow` in the stats row and makes
    # results() raise (the kernel has substituted frame 0, which is not what the reference would have drawn).
    reseed_pool: Optional[int] = None
    compute_centroids: bool = True
    gather_events: bool = False  # also materialise the event-major packed token copy (csb_embed_scatter_ordered makes it unnecessary)
    # phase 2 without an early pixel mask: run the pixel-difference kernel on the side stream next to the selection launch
    # (the Exp(1) / pool draws of phase 1 always run next to the Gram kernel)
    parallel_branches: bool = False
    # phase1(pixel_values=...) starts the raw pixel-difference mask on a side stream right after the Gram kernel, next to the
    # k-means solver.  Measured at config 2 (scripts/overlap_probe2.py): with L2 evict-first pixel loads and the grid capped at
    # `early_pixel_ctas` CTAs the 508 MB pass (160 us on 64 SMs) hides completely under the ~190 us latency-bound solver without
    # slowing it (solver + pixel back to back 321 us, overlapped 207 us); an uncapped grid stretches the solver by ~15 us.
    early_pixel_ctas: int = int(os.environ.get("CSB_EARLY_PIX_CTAS", "64"))
    # randomness of the clustering (first centre, Exp(1) sampling noise, re-seed indices): True = the solver's built-in Philox
    # generator (csb_kmeans_params.rng_state: nothing is launched in front of the solver, the re-seed supply cannot run out);
    # False = torch's CUDA generator (3 small kernels per stream).  Injected init_indices / exp_noise / reseed_pool always win.
    device_rng: bool = os.environ.get("CSB_DEVICE_RNG", "1") != "0"
    # with device_rng=False: run torch's draws on the side stream next to the Gram kernel (-28 us for a stream that runs alone,
    # -5 % throughput with six streams in flight: more graph branches to schedule)
    parallel_draws: bool = os.environ.get("CSB_PAR_DRAWS", "0") != "0"
    # sequential frame passes per warp in the solver's assignment pass (csb_kmeans_params.frames_per_warp); 0 = the library's
    # choice (k <= 64: frames side by side on sub-warp groups, one pass; else two passes).  More passes = fewer co-resident CTAs
    solver_frames_per_warp: int = 0

    def clusters_for(self, T: int) -> int:
        if self.cluster_num is not None:
            return self.cluster_num
        return math.ceil(T / self.frames_per_event)

    def gate_open(self, T: int) -> bool:
        k = self.clusters_for(T)
        if T <= k:
            return False
        return self.cluster_num is not None or k > self.min_events


class StreamCompressor:
    """Buffers + launch sequence for streams of one shape (T frames, P tokens, D channels, W pixel width)."""

    def __init__(self, T: int, P: int, D: int, W: int, n_ids: int, *, device, cfg: Optional[StreamConfig] = None,
                 ts_dtype=torch.float64, embed_dim: int = 3584, embed_dtype=torch.bfloat16, image_token_index: int = 151665,
                 items: Optional[Sequence] = None):
        """``items``: the visual items of the prompt in order, ``[(n_frames, modal), ...]`` with modal "video" / "image"
        (cogreasoner_chat.py:392-430 iterates over them); default = one video item of T frames.  All items share P tokens per
        frame (the reference's ``features.view(total_image_num, P, D)``, :279, needs that too).  Every video item is its own
        pixel-difference segment (its first frame is fully kept); image items and single-frame videos keep every token."""
        self.cfg = cfg or StreamConfig()
        self.T, self.P, self.D, self.W, self.n_ids = T, P, D, W, n_ids
        self.K = P * D
        self.k = self.cfg.clusters_for(T)
        self.device = torch.device(device)
        self.image_token_index = image_token_index
        self.gate = self.cfg.gate_open(T)
        dev = self.device
        ops.check_device()
        k = self.k
        if self.gate:
            self.G = torch.empty(T, T, dtype=torch.float64, device=dev)
            self.gram_ws = torch.empty(ops.gram_workspace_bytes(T, self.K), dtype=torch.uint8, device=dev)
            self.km = ops.KMeansBuffers(T, k, ts_dtype, dev)
            self.noise = torch.empty(max(k - 1, 1), T, dtype=torch.float32, device=dev) if not self.cfg.device_rng else None
            # {seed, offset} of the solver's built-in generator, seeded from torch's CUDA generator (torch.manual_seed reproduces it)
            self.rng_state = torch.randint(0, 2 ** 62, (2,), dtype=torch.int64, device=dev)
            self.rng_state[1] = 0
            n_pool = self.cfg.reseed_pool if self.cfg.reseed_pool is not None else k * (self.cfg.max_iteration + 1)
            self.pool = torch.zeros(max(1, n_pool), dtype=torch.int64, device=dev)
            self.centroids = torch.empty(k, self.K, dtype=torch.float32, device=dev) if self.cfg.compute_centroids else None
            self.packed = torch.empty(T, self.K, dtype=torch.bfloat16, device=dev) if self.cfg.gather_events else None
        self.items = [(int(t), str(m)) for t, m in items] if items is not None else [(T, "video")]
        if sum(t for t, _ in self.items) != T or any(t < 1 or m not in ("video", "image") for t, m in self.items):
            raise ValueError(f"items {self.items} do not add up to T = {T} frames of modal 'video' / 'image'")
        kind = torch.zeros(T, dtype=torch.uint8)
        self._segments = []  # (first frame, frames) of the video items that go through the pixel-difference kernel
        f0 = 0
        for t, m in self.items:
            if m == "image" or t == 1:  # :401-403
                kind[f0: f0 + t] = 1
            else:
                self._segments.append((f0, t))
            f0 += t
        self.frame_kind = kind.to(dev) if bool(kind.any()) else None
        self._row_base = torch.tensor([T, self.k if self.gate else 0, 0, 0, 0, 0, 0, 0, 0], dtype=torch.int64, device=dev)
        self._ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        self.minor_flags = torch.zeros(T, dtype=torch.uint8, device=dev)
        self.n_minor = torch.zeros(1, dtype=torch.int64, device=dev)
        self.mask = torch.empty(T * P, dtype=torch.uint8, device=dev)
        self.out = torch.empty(T * P, D, dtype=torch.bfloat16, device=dev)
        self.compact_ws = ops.compact_stream_workspace(T * P, n_ids, dev)  # zero-filled: carries the scan kernel's barrier epoch
        self._sel, self._cmp = {}, {}  # output buffers of the selection / compaction launches, allocated on first use
        # the raw pixel-difference mask is independent of the clustering: phase1 may launch it on a side stream so that
        # this HBM-bound kernel overlaps the latency-bound k-means solver
        self.side = torch.cuda.Stream(device=dev)
        self._ev_fork = torch.cuda.Event()
        self._ev_diff = torch.cuda.Event()
        self._ev_branch = torch.cuda.Event()
        self._ev_join = torch.cuda.Event()
        self._diff_pending = False
        self.order = self.ev_off = self.dist = self.reps = self.rep_counts = self.rep_flags = None
        self.cos = self.low = self.minor_list = self.events_repr = None
        self.kept_index = self.n_kept = None
        self.out_ids = self.out_attention = self.n_ids_out = self.n_image_ids = None
        self._G_used = None

    # ------------------------------------------------------------------------------------------
    def _branch(self):
        """Context manager: run the body on the side stream, ordered after everything enqueued so far on the current stream;
        ``_join()`` makes the current stream wait for it.  Inside a graph capture this becomes a parallel branch."""
        cur = torch.cuda.current_stream(self.device)
        self._ev_branch.record(cur)
        self.side.wait_event(self._ev_branch)
        return torch.cuda.stream(self.side)

    def _join(self):
        self._ev_join.record(self.side)
        torch.cuda.current_stream(self.device).wait_event(self._ev_join)

    def _pixel_diff(self, pixel_values: torch.Tensor, stream=None, max_ctas: int = 0) -> None:
        """Raw pixel-difference mask of every video item (each item is its own segment, :405-412) into ``self.mask``; the rows of
        image-like frames are filled by the compaction kernel (``frame_kind``)."""
        P, W = self.P, self.W
        rows_per_frame = pixel_values.shape[0] // self.T
        for f0, t in self._segments:
            ops.pixel_diff(pixel_values[f0 * rows_per_frame: (f0 + t) * rows_per_frame], t, P, W, threshold=self.cfg.pixel_threshold,
                           out=self.mask[f0 * P: (f0 + t) * P], stream=stream, max_ctas=max_ctas)

    # ------------------------------------------------------------------------------------------
    def phase1(self, features: torch.Tensor, timestamps: torch.Tensor, *, init_indices: Optional[torch.Tensor] = None,
               first_index: Optional[int] = None, exp_noise: Optional[torch.Tensor] = None,
               reseed_pool: Optional[torch.Tensor] = None, pixel_values: Optional[torch.Tensor] = None,
               gram: Optional[torch.Tensor] = None) -> None:
        """features [T*P, D] bf16, timestamps [T] on the device.  Enqueues everything; returns at once.
        ``gram`` (optional): a ready [T,T] fp64 Gram matrix of these frames (e.g. ``StreamCache.gram``, kept up to date
        incrementally); the Gram kernel is then skipped.
        ``pixel_values`` (optional, the same tensor phase2 gets): start the raw pixel-difference mask now, on the side stream,
        overlapped with the k-means solver (throttled grid + L2 evict-first loads, see ``StreamConfig.early_pixel_ctas``); the
        mask does not depend on the clustering or the question.  phase1 joins the side stream before it returns."""
        if not self.gate:
            return
        T, k = self.T, self.k
        x2 = features.view(T, self.K)
        fork = pixel_values is not None

        def draws():
            """-> dict(n_given, noise, pool, rng).  Default: the solver's built-in generator (no launch at all); injected values and
            ``device_rng=False`` go through the caller's / torch's generators as before."""
            out = dict(noise=None, pool=reseed_pool, rng=None)
            use_rng = self.cfg.device_rng and exp_noise is None
            if use_rng:
                out["rng"] = self.rng_state
            elif reseed_pool is None:
                out["pool"] = self.pool
                self.pool.random_(0, T)
            if init_indices is not None:
                self.km.init_indices.copy_(init_indices, non_blocking=True)
                out["n_given"] = k
                return out
            if use_rng and (first_index is None or first_index == "device"):
                out["n_given"] = 0  # the first centre comes from the built-in generator too (kmeans_with_time.py:41)
                return out
            if first_index == "device":  # drawn by the CUDA generator: no host value baked into a captured graph
                torch.randint(0, T, (1,), device=self.device, out=self.km.init_indices[:1])
            else:
                first = random.randint(0, T - 1) if first_index is None else first_index  # kmeans_with_time.py:41
                self.km.init_indices[:1].fill_(first)
            out["n_given"] = 1
            if exp_noise is not None:
                out["noise"] = exp_noise
            elif not use_rng and k > 1:
                nz = self.noise[: k - 1]
                nz.exponential_(1)  # one batched draw (the drop-in function draws per round like ATen does)
                out["noise"] = nz
            return out

        G = self.G if gram is None else gram
        if self.cfg.parallel_draws and not self.cfg.device_rng:
            # torch-generator draws (three tiny kernels) do not depend on the Gram matrix: they run on the side stream next to it
            with self._branch():
                d_ = draws()
            if gram is None:
                ops.gram(x2, impl=self.cfg.gram_impl, out=self.G, workspace=self.gram_ws)
            if fork:
                self._ev_fork.record(torch.cuda.current_stream(self.device))
            self._join()
        else:
            if gram is None:
                ops.gram(x2, impl=self.cfg.gram_impl, out=self.G, workspace=self.gram_ws)
            if fork:
                self._ev_fork.record(torch.cuda.current_stream(self.device))
            d_ = draws()
        self._G_used = G
        ops.kmeans_solve(G, timestamps, self.km, n_init_given=d_["n_given"], exp_noise=d_["noise"] if k > 1 else None,
                         reseed_pool=d_["pool"], alpha=self.cfg.alpha, max_iteration=self.cfg.max_iteration, tol=self.cfg.tol,
                         frames_per_warp=self.cfg.solver_frames_per_warp, rng_state=d_["rng"])
        if fork:
            # enqueued AFTER the cooperative solver so its CTAs become resident first; the HBM-bound difference kernel then
            # fills the rest of the machine while the (latency-bound) solver iterates
            self.side.wait_event(self._ev_fork)
            self._pixel_diff(pixel_values, stream=self.side, max_ctas=self.cfg.early_pixel_ctas)
            self._ev_diff.record(self.side)
            self._diff_pending = True
        self.order, self.ev_off = self.km.event_order, self.km.event_offsets
        # one pass over X: the fp32 centroids the reference returns (kmeans_with_time.py:134); the chronological per-event
        # token gather (cogreasoner_chat.py:291-296) is NOT materialised: csb_embed_scatter_ordered reads the feature rows
        # through `event_order` when the summary prompts are embedded (cfg.gather_events=True restores the packed copy)
        if self.centroids is not None or self.packed is not None:
            ops.centroid_gather(x2, self.km.member_offsets, self.km.members, self.km.event_offsets, k, self.centroids, self.packed)
        if fork:  # join: everything phase 1 started is ordered before whatever follows on this stream (and a capture can end)
            torch.cuda.current_stream(self.device).wait_event(self._ev_diff)

    # ------------------------------------------------------------------------------------------
    def _select(self, events_repr, question, hidden, cu_seqlens):
        """cogreasoner_chat.py:284 + :317-318 + :325-333: representatives, relevance flags, minor frames."""
        T, k = self.T, self.k
        G = self._G_used
        q = question.view(-1)
        if T <= 256 * k:  # one fused launch (a warp per event for the representatives, quadratic in the event size)
            b = ops.relevance_select(q, events_repr, G, self.ev_off, self.order, self.km.assignments, self._ticket, k=k,
                                     n_rep=self.cfg.n_representatives, threshold=self.cfg.relevance_threshold, hidden=hidden,
                                     cu_seqlens=cu_seqlens, bufs=self._sel)
            self.dist, self.reps, self.rep_counts, self.rep_flags = b["dist"], b["reps"], b["rep_counts"], b["rep_flags"]
            self.cos, self.low, self.minor_flags, self.minor_list, self.n_minor = b["cos"], b["low"], b["minor_flags"], b["minor_list"], b["n_minor"]
            if events_repr is None:
                self.events_repr = b["events"]
            return
        if events_repr is None:
            events_repr = ops.seq_mean_pool(hidden, cu_seqlens)
            self.events_repr = events_repr
        assign = self.km.assignments
        self.dist = ops.dist_to_assigned_gram(G, assign, self.ev_off, self.order, k)
        self.reps, self.rep_counts, self.rep_flags = ops.select_representatives(self.dist, assign, k, self.cfg.n_representatives)
        self.cos, self.low, self.minor_flags, self.minor_list, self.n_minor = ops.relevance_minor(
            q, events_repr, assign, self.rep_flags, self._ticket, self.cfg.relevance_threshold)

    def phase2(self, features: torch.Tensor, events_repr: Optional[torch.Tensor], question: Optional[torch.Tensor],
               pixel_values: torch.Tensor, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, *,
               hidden: Optional[torch.Tensor] = None, cu_seqlens: Optional[torch.Tensor] = None) -> None:
        """events_repr [k, Dl], question [1, Dl] (frozen-LLM outputs); pixel_values [T*P*m*m, C]; ids [n_ids].
        Instead of ``events_repr`` the packed last hidden states of the k summary forwards can be given (``hidden``
        [sum(len), Dl] + ``cu_seqlens`` [k+1] int32): their mean pool (:317-318) is then fused into the selection launch."""
        T, P = self.T, self.P
        par = self.cfg.parallel_branches and self.gate and not self._diff_pending
        if par:  # the raw pixel-difference mask does not depend on the relevance test: a parallel branch
            with self._branch():
                self._pixel_diff(pixel_values)
        if self.gate:
            self._select(events_repr, question, hidden, cu_seqlens)
        else:
            self.minor_flags.zero_()
            self.n_minor.zero_()
        if par:
            self._join()
        elif self._diff_pending:  # phase 1 computed (and joined) the raw mask of these pixels
            self._diff_pending = False
        else:
            self._pixel_diff(pixel_values)
        # per-frame mask rules + scan + compaction (minor frames emitted as their pooled token) + id filter
        b = ops.compact_stream(features, T, P, self.mask, self.minor_flags, min_tokens=self.cfg.min_tokens, input_ids=input_ids,
                               image_token_index=self.image_token_index, attention_mask=attention_mask, out=self.out,
                               workspace=self.compact_ws, bufs=self._cmp, frame_kind=self.frame_kind)
        self.kept_index, self.n_kept = b["kept_index"], b["n_kept"]
        self.out_ids, self.out_attention = b["out_ids"], b.get("out_attention")
        self.n_ids_out, self.n_image_ids = b["n_out"][0:1], b["n_out"][1:2]

    # ------------------------------------------------------------------------------------------
    def capture(self, features, timestamps, events_repr, question, pixel_values, input_ids, attention_mask=None, *,
                early_pixel: bool = True):
        """Capture both phases into CUDA graphs bound to these input tensors (refill them in place, then ``replay1()`` /
        ``replay2()``).  The first centre is drawn on the device; everything else is identical to the eager path.
        ``early_pixel``: the raw pixel-difference mask is a parallel branch of the phase-1 graph (next to the solver)."""
        pv1 = pixel_values if (early_pixel and self.gate) else None
        cur = torch.cuda.current_stream(self.device)
        warm = torch.cuda.Stream(device=self.device)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):  # one eager pass: lazy initialisation must not happen inside the capture
            self.phase1(features, timestamps, first_index="device", pixel_values=pv1)
            self.phase2(features, events_repr, question, pixel_values, input_ids, attention_mask)
        cur.wait_stream(warm)
        torch.cuda.synchronize(self.device)
        self.graph1, self.graph2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph1):
            self.phase1(features, timestamps, first_index="device", pixel_values=pv1)
        with torch.cuda.graph(self.graph2, pool=self.graph1.pool()):
            self._diff_pending = pv1 is not None  # graph 1 always leaves the raw mask behind
            self.phase2(features, events_repr, question, pixel_values, input_ids, attention_mask)
            self._stats_static = self.stats_row()
        return self

    def replay1(self):
        self.graph1.replay()

    def replay2(self):
        self.graph2.replay()

    # ------------------------------------------------------------------------------------------
    def stats_row(self) -> torch.Tensor:
        """[len(STATS_FIELDS)] int64 on the device (no sync, capture-safe): the per-stream row that ranks all-gather."""
        row = self._row_base.clone()
        if self.gate:
            st = self.km.status
            row[2:3] = st[0:1]
            row[3:4] = st[2:3]
            row[7:8] = st[1:2]
            row[8:9] = st[4:5]
        row[4:5] = self.n_minor
        row[5:6] = self.n_kept
        row[6:7] = self.n_ids_out
        return row

    def results(self) -> Dict[str, object]:
        """Host view of the outputs (synchronises)."""
        nk = int(self.n_kept.item())
        no = int(self.n_ids_out.item())
        n_img = int(self.n_image_ids.item())
        if n_img != self.T * self.P:  # the reference's `text_masks[image_selected] = compression_mask` raises here (:463)
            raise RuntimeError(f"shape mismatch: {n_img} <image> ids for a compression mask of {self.T * self.P} tokens")
        res = dict(n_kept=nk, mask=self.mask.view(torch.bool), features=self.out[:nk], kept_index=self.kept_index[:nk],
                   input_ids=self.out_ids[:no],
                   attention_mask=None if self.out_attention is None else self.out_attention[:no])
        if self.gate:
            nm = int(self.n_minor.item())
            cnt = self.rep_counts.tolist()
            status = ops.decode_status(self.km.status)
            if status["reseed_overflow"]:
                raise RuntimeError(
                    f"StreamCompressor: the re-seed pool ran out ({status['n_reseed_used']} draws needed); the clustering is not the "
                    "reference's.  Pass a larger reseed_pool (StreamConfig.reseed_pool=None sizes it to the upper bound)")
            res.update(
                status=status, assignments=self.km.assignments, init_indices=self.km.init_indices,
                centroid_times=self.km.centroid_times,
                centroids=None if self.centroids is None else self.centroids.view(self.k, self.P, self.D),
                representatives=[self.reps[i, : cnt[i]] for i in range(self.k)], minor=self.minor_list[:nm].tolist(),
                cos=self.cos, low_flags=self.low, event_order=self.order, event_offsets=self.ev_off, packed_events=self.packed,
            )
        else:
            res.update(minor=[])
        return res


# ---------------------------------------------------------------------------------------------------
# multi-GPU: independent streams, one stats all-gather
# ---------------------------------------------------------------------------------------------------
def shard_streams(n_streams: int, rank: int, world_size: int) -> List[int]:
    """Stream s runs on rank s mod W (DistributedSampler-style round robin, answer_generate.py:186)."""
    return list(range(rank, n_streams, world_size))


def gather_stats(local_rows: torch.Tensor, world_size: Optional[int] = None) -> torch.Tensor:
    """All-gather the [n_local, len(STATS_FIELDS)] int64 stats rows -> [world, n_local, F].

    The only collective on the path (NCCL over NVLink on GPUs, gloo in the CPU tests).  Every rank must
    pass the same number of rows (pad with zero rows when the stream count does not divide).
    """
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local_rows.unsqueeze(0)
    ws = world_size or dist.get_world_size()
    rows = local_rows.contiguous()
    out = torch.empty((ws * rows.shape[0],) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(out, rows)  # rank-major concatenation along dim 0
    return out.view((ws,) + tuple(rows.shape))


def pad_rows(rows: Sequence[torch.Tensor], n: int, device) -> torch.Tensor:
    """Stack per-stream rows and pad with zero rows to ``n`` (ranks with fewer streams)."""
    F = len(STATS_FIELDS)
    out = torch.zeros(n, F, dtype=torch.int64, device=device)
    for i, r in enumerate(rows):
        out[i] = r
    return out
