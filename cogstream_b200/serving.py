"""Many independent streams on one GPU, and over the GPUs of one box (BASELINE config 3).

The reference serves videos share-nothing: one process per GPU, a ``DistributedSampler`` over videos, no collective
(``evaluate/answer_generate.py:154-187``).  Here a rank keeps ``depth`` streams in flight on separate CUDA streams, each with
its own :class:`~cogstream_b200.pipeline.StreamCompressor` workspaces (a stream's latency-bound k-means solver overlaps the
bandwidth-bound kernels of the others), and the ranks exchange exactly one collective: the all-gather of the per-stream stats
rows.  Streams of a batch must share a shape ``(T, P, D, W, n_ids)``; one :class:`StreamPool` per shape.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from .pipeline import STATS_FIELDS, StreamCompressor, StreamConfig, gather_stats, pad_rows, shard_streams


@dataclass
class StreamInputs:
    """Device (or pinned-host) tensors of one stream, in the layouts the reference hands to the hot path."""

    features: torch.Tensor  # [T*P, D] bf16
    timestamps: torch.Tensor  # [T] f64 / f32
    pixel_values: torch.Tensor  # [T*P*m*m, C] bf16
    input_ids: torch.Tensor  # [n_ids] int64
    attention_mask: Optional[torch.Tensor] = None


# (events [k*?, D] packed view, event_offsets, timestamps) -> (events_repr [k, Dl], question [1, Dl]): the frozen LLM
EmbedFn = Callable[[StreamCompressor, StreamInputs], Tuple[torch.Tensor, torch.Tensor]]


class _Lane:
    def __init__(self, T, P, D, W, n_ids, device, cfg, ts_dtype):
        self.stream = torch.cuda.Stream(device=device)
        self.sc = StreamCompressor(T, P, D, W, n_ids, device=device, cfg=cfg, ts_dtype=ts_dtype)
        self.busy_with: Optional[int] = None
        self.staging: Optional[StreamInputs] = None

    def stage(self, like: StreamInputs, device) -> StreamInputs:
        if self.staging is None:
            def dev(t):
                return None if t is None else torch.empty(t.shape, dtype=t.dtype, device=device)
            self.staging = StreamInputs(dev(like.features), dev(like.timestamps), dev(like.pixel_values), dev(like.input_ids),
                                        dev(like.attention_mask))
        return self.staging


class StreamPool:
    """``depth`` compressors of one stream shape, used round-robin."""

    def __init__(self, T: int, P: int, D: int, W: int, n_ids: int, *, device, depth: int = 4, cfg: Optional[StreamConfig] = None,
                 ts_dtype=torch.float64):
        self.device = torch.device(device)
        self.shape = (T, P, D, W, n_ids)
        self.lanes = [_Lane(T, P, D, W, n_ids, self.device, cfg, ts_dtype) for _ in range(max(1, depth))]

    # ------------------------------------------------------------------------------------------------
    def run_batch(self, streams: Sequence[StreamInputs], embed: EmbedFn, *, keep_outputs: bool = True,
                  kmeans_kw: Optional[Sequence[dict]] = None) -> Tuple[List[Optional[Dict[str, object]]], torch.Tensor]:
        """Run every stream through phase 1 -> ``embed`` (the frozen LLM, or a stand-in) -> phase 2.

        Host (pinned) inputs are copied to per-lane staging buffers on the lane's stream; device inputs are used in place.
        Returns ``(results, stats_rows [n, len(STATS_FIELDS)] int64 on the device)``; ``results[i]`` is
        ``StreamCompressor.results()`` of stream i (cloned, so the lane can be reused) or ``None`` when
        ``keep_outputs=False`` (throughput runs: only the stats rows are kept; their ``reseed_overflow`` column is non-zero
        for a stream whose caller-supplied re-seed pool ran out -- ``check_rows`` raises on it).
        """
        n = len(streams)
        main = torch.cuda.current_stream(self.device)
        results: List[Optional[Dict[str, object]]] = [None] * n
        # allocated and zero-filled on `main` BEFORE the start event: the lanes (which wait on it) write their rows after the fill
        rows = torch.zeros(n, len(STATS_FIELDS), dtype=torch.int64, device=self.device)
        start = torch.cuda.Event()
        start.record(main)
        for L in self.lanes:
            L.stream.wait_event(start)
            rows.record_stream(L.stream)
            L.busy_with = None

        def finish(L: _Lane):
            i = L.busy_with
            if i is None:
                return
            if keep_outputs:
                L.stream.synchronize()
                res = L.sc.results()
                results[i] = {k: (v.clone() if isinstance(v, torch.Tensor) else
                                  [t.clone() for t in v] if isinstance(v, list) and v and isinstance(v[0], torch.Tensor) else v)
                              for k, v in res.items()}
            L.busy_with = None

        for i, s in enumerate(streams):
            L = self.lanes[i % len(self.lanes)]
            finish(L)
            with torch.cuda.stream(L.stream):
                x = s
                if not s.features.is_cuda:
                    st = L.stage(s, self.device)
                    st.features.copy_(s.features, non_blocking=True)
                    st.timestamps.copy_(s.timestamps, non_blocking=True)
                    st.pixel_values.copy_(s.pixel_values, non_blocking=True)
                    st.input_ids.copy_(s.input_ids, non_blocking=True)
                    if s.attention_mask is not None:
                        st.attention_mask.copy_(s.attention_mask, non_blocking=True)
                    x = st
                kw = kmeans_kw[i] if kmeans_kw is not None else {}
                L.sc.phase1(x.features, x.timestamps, **kw)
                ev, q = embed(L.sc, x) if L.sc.gate else (None, None)
                L.sc.phase2(x.features, ev, q, x.pixel_values, x.input_ids, x.attention_mask)
                rows[i].copy_(L.sc.stats_row(), non_blocking=True)
            L.busy_with = i
        for L in self.lanes:
            finish(L)
            done = torch.cuda.Event()
            done.record(L.stream)
            main.wait_event(done)
        return results, rows


def check_rows(rows: torch.Tensor) -> torch.Tensor:
    """Raise if any stats row (``[..., len(STATS_FIELDS)]``) reports an exhausted re-seed pool (synchronises)."""
    col = STATS_FIELDS.index("reseed_overflow")
    bad = torch.nonzero(rows[..., col].reshape(-1))[:, 0].tolist()
    if bad:
        raise RuntimeError(f"re-seed pool exhausted for stream rows {bad}: their clustering is not the reference's; rerun them "
                           "with a larger reseed_pool")
    return rows


def run_sharded(pool: StreamPool, all_streams: Sequence[StreamInputs], embed: EmbedFn, rank: int, world_size: int, *,
                keep_outputs: bool = False):
    """Rank ``rank`` of ``world_size`` processes streams ``rank, rank + W, ...``; one all-gather of the stats rows follows.

    Returns ``(local_results, stats [world_size, per_rank, len(STATS_FIELDS)])``; rows of ranks with fewer streams are zero.
    """
    mine = shard_streams(len(all_streams), rank, world_size)
    res, rows = pool.run_batch([all_streams[i] for i in mine], embed, keep_outputs=keep_outputs)
    per_rank = (len(all_streams) + world_size - 1) // world_size
    padded = pad_rows(list(rows), per_rank, rows.device)
    return dict(zip(mine, res)), check_rows(gather_stats(padded, world_size))
