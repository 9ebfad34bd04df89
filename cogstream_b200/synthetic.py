"""Synthetic stream generator.

Everything is produced on the CPU from a seeded ``torch.Generator`` so the CPU oracle and
the CUDA path see bit-identical inputs; callers move the tensors to the device.

Shapes follow what ``model/processing_cogreasoner.py`` / ``model/cogreasoner_chat.py`` of
the reference hand to the hot path:
  features     [T*P, D] bf16          projector output (``encode_images``, cogreasoner_chat.py:264-276)
  timestamps   [T] float64            ``self.all_timestamps`` (cogreasoner_chat.py:816, processing_cogreasoner.py:487-507)
  pixel_values [T*P*m*m, C] bf16      merge-block-contiguous patches (image_processing_videollama3.py:326-345)
  grid_sizes   [[T, gh, gw]] int64, merge_sizes [m] int64, modals ["video"]
  input_ids    [n_text_head + T*P + n_text_tail] int64 with ``image_token_index`` at visual slots
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch

IMAGE_TOKEN_INDEX = 151665  # model/config.json image_token_index


@dataclass
class SyntheticStream:
    T: int
    P: int
    D: int
    features: torch.Tensor  # [T*P, D] bf16
    timestamps: torch.Tensor  # [T] f64 (or f32)
    pixel_values: torch.Tensor  # [T*P*m*m, C] bf16
    grid_sizes: torch.Tensor  # [1,3] int64
    merge_sizes: torch.Tensor  # [1] int64
    modals: List[str]
    input_ids: torch.Tensor  # [N] int64
    attention_mask: torch.Tensor  # [N] int64
    scene_of_frame: torch.Tensor  # [T] int64 (ground truth, diagnostics only)

    @property
    def k(self) -> int:
        return math.ceil(self.T / 15)

    def to(self, device) -> "SyntheticStream":
        kw = {}
        for name, v in self.__dict__.items():
            kw[name] = v.to(device, non_blocking=True) if isinstance(v, torch.Tensor) else v
        return SyntheticStream(**kw)

    def pin(self) -> "SyntheticStream":
        kw = {}
        for name, v in self.__dict__.items():
            kw[name] = v.pin_memory() if isinstance(v, torch.Tensor) else v
        return SyntheticStream(**kw)


def _grid_for_tokens(P: int):
    """Pick (gh, gw) in merged tokens with gh*gw == P, as square as possible."""
    gh = int(math.isqrt(P))
    while P % gh:
        gh -= 1
    return gh, P // gh


def make_features(T: int, P: int, D: int, seed: int = 0, *, iid: bool = False, noise: float = 0.3,
                  chunk_frames: int = 64):
    """Scene-structured features: ``S = max(2, T//20)`` prototypes ``N(0,1) [S,P,D]``, ``S-1``
    random cut points, ``X[t] = proto[scene(t)] + noise * N(0,1)``, cast to bf16.
    ``iid=True`` gives plain ``N(0,1)`` (stress case).  Returns ``([T*P, D] bf16, scene [T])``.
    """
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(T, P, D, dtype=torch.bfloat16)
    if iid:
        scene = torch.zeros(T, dtype=torch.long)
        for s in range(0, T, chunk_frames):
            e = min(T, s + chunk_frames)
            out[s:e] = torch.randn(e - s, P, D, generator=g).to(torch.bfloat16)
        return out.view(T * P, D), scene
    S = max(2, T // 20)
    S = min(S, T)
    cuts = torch.randperm(T - 1, generator=g)[: S - 1].add(1).sort().values
    scene = torch.bucketize(torch.arange(T), cuts, right=True)
    protos = torch.randn(S, P, D, generator=g)
    for s in range(0, T, chunk_frames):
        e = min(T, s + chunk_frames)
        x = protos[scene[s:e]] + noise * torch.randn(e - s, P, D, generator=g)
        out[s:e] = x.to(torch.bfloat16)
    return out.view(T * P, D), scene


def make_pixels(T: int, P: int, merge: int = 2, C: int = 588, seed: int = 0, *, change_frac: float = 0.2,
                static_frame_frac: float = 0.05, near_threshold_frac: float = 0.0):
    """``U(-1,1)`` pixels, frame f = frame f-1 with ``change_frac`` of its merged tokens
    re-drawn (so the pixel-difference mask keeps about that share); a few frames are exact
    copies of their predecessor (exercises the ``min_tokens`` rule, cogreasoner_chat.py:413-414).
    ``near_threshold_frac`` of tokens get a tiny perturbation so mean|diff|*255 lands near 0.1.
    Returns ``[T*P*merge*merge, C] bf16``.
    """
    g = torch.Generator().manual_seed(seed + 7919)
    W = merge * merge * C
    pix = torch.empty(T, P, W, dtype=torch.bfloat16)
    cur = torch.rand(P, W, generator=g).mul_(2).sub_(1).to(torch.bfloat16)
    pix[0] = cur
    for f in range(1, T):
        cur = cur.clone()
        if torch.rand((), generator=g).item() >= static_frame_frac:
            changed = torch.rand(P, generator=g) < change_frac
            n = int(changed.sum())
            if n:
                cur[changed] = torch.rand(n, W, generator=g).mul_(2).sub_(1).to(torch.bfloat16)
            if near_threshold_frac > 0:
                near = (torch.rand(P, generator=g) < near_threshold_frac) & ~changed
                n = int(near.sum())
                if n:
                    # flip the sign of a handful of elements: mean|diff|*255 ~ 0.05 .. 0.3
                    sel = cur[near].float()
                    cols = torch.randint(0, W, (n, 3), generator=g)
                    eps = torch.rand(n, 3, generator=g) * (0.4 * W / (3 * 255))
                    rows = torch.arange(n).unsqueeze(1).expand_as(cols)
                    sel[rows, cols] = sel[rows, cols] + eps
                    cur[near] = sel.to(torch.bfloat16)
        pix[f] = cur
    return pix.view(T * P * merge * merge, C)


def make_embeddings(k: int, Dl: int = 3584, seed: int = 0, *, low_frac: float = 0.5, dtype=torch.bfloat16):
    """Question embedding ``N(0,1) [1,Dl]`` and event embeddings mixed so that about
    ``low_frac`` of the events have cosine below the 0.45 threshold (cogreasoner_chat.py:329).
    Returns ``(events_repr [k,Dl], question [1,Dl])`` in ``dtype``.
    """
    g = torch.Generator().manual_seed(seed + 104729)
    q = torch.randn(1, Dl, generator=g)
    noise = torch.randn(k, Dl, generator=g)
    target = torch.where(torch.rand(k, generator=g) < low_frac,
                         0.05 + 0.35 * torch.rand(k, generator=g),
                         0.50 + 0.40 * torch.rand(k, generator=g))
    # cos(a*q + b*n, q) ~= a / sqrt(a^2 + b^2) for independent q, n
    a = target
    b = torch.sqrt(1 - target ** 2)
    ev = a.unsqueeze(1) * q + b.unsqueeze(1) * noise
    return ev.to(dtype), q.to(dtype)


def make_stream(T: int, P: int, D: int, seed: int = 0, *, fps: float = 1.0, merge: int = 2, C: int = 588,
                iid: bool = False, ts_dtype=torch.float64, n_text_head: int = 14, n_text_tail: int = 23,
                change_frac: float = 0.2, near_threshold_frac: float = 0.0, items: Optional[List] = None) -> SyntheticStream:
    """``items``: optional ``[(n_frames, modal), ...]`` split of the T frames into visual items (several video segments, images,
    single-frame videos: what ``grid_sizes`` / ``modals`` carry for multi-item prompts, cogreasoner_chat.py:392-430)."""
    feats, scene = make_features(T, P, D, seed, iid=iid)
    ts = (torch.arange(T, dtype=torch.float64) / fps).to(ts_dtype)
    pix = make_pixels(T, P, merge, C, seed, change_frac=change_frac, near_threshold_frac=near_threshold_frac)
    gh, gw = _grid_for_tokens(P)
    if items is None:
        items = [(T, "video")]
    assert sum(t for t, _ in items) == T
    grid = torch.tensor([[t, gh * merge, gw * merge] for t, _ in items], dtype=torch.long)
    ids = torch.cat([
        torch.arange(1000, 1000 + n_text_head, dtype=torch.long),
        torch.full((T * P,), IMAGE_TOKEN_INDEX, dtype=torch.long),
        torch.arange(2000, 2000 + n_text_tail, dtype=torch.long),
    ])
    return SyntheticStream(T, P, D, feats, ts, pix, grid, torch.tensor([merge] * len(items), dtype=torch.long),
                           [m for _, m in items], ids, torch.ones_like(ids), scene)
