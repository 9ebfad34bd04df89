"""CPU oracle for CogReasoner's Visual Stream Compression hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``cogstream_b200/`` may import this file; the only
allowed users are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` (as the checker / the CPU arm, never as the
product path).

This is a *restatement* (not a copy) of the reference's algorithm, written against
torch-CPU tensor ops because the reference's arithmetic *is* ATen arithmetic
(``torch.cdist``'s augmented sgemm, bf16 op-by-op rounding, ``multinomial`` = argmax(p/Exp(1))).
The reference is pure Python + torch (no C/C++ to compile into ``oracle/_ref``); it is
imported live from ``/root/reference`` by ``oracle/make_golden.py`` to (1) validate this
restatement bit-for-bit and (2) produce the committed fixtures in ``tests/golden/``.
The reference ships no tests / golden vectors of its own, so parity
is pinned by those generated fixtures: outputs of the unmodified reference run in the
build container (torch 2.11.0 CPU).

Differences to the reference that are deliberate and result-neutral:
  * the per-row Python loop of ``kmeans_with_time.py:81-101`` is evaluated for all rows at
    once with the same elementwise fp ops (bit-identical, much faster);
  * RNG draws can be *injected* (``init_indices``, ``exp_noise``, ``reseed_indices``) so
    that "same inputs and same init" is well defined across CPU and GPU generators.
    Without injection the same host RNG calls as the reference are made, in the same order.

Reference citations are ``file:line`` relative to ``/root/reference``.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

__all__ = [
    "KMeansInfo",
    "kmeans_with_time_min_max",
    "select_additional_frames",
    "event_gather",
    "cosine_low_events",
    "minor_frames",
    "compress_unimportant_events",
    "get_compression_mask",
    "compress_visual_tokens",
    "encoder_tail",
    "phase1",
    "phase2",
    "Phase1Result",
    "Phase2Result",
]


@dataclass
class KMeansInfo:
    """Side information the reference does not return but parity tests need."""

    init_indices: List[int] = field(default_factory=list)
    reseed_indices: List[int] = field(default_factory=list)  # consumed, in order
    iterations: int = 0
    converged: bool = False
    degenerate_rounds: int = 0  # init rounds that took the probs_sum == 0 branch
    min_margin: float = float("inf")  # smallest (2nd best - best) final distance seen


def _row_minmax_normalise(d: torch.Tensor) -> torch.Tensor:
    """model/kmeans_with_time.py:86-99, all rows at once.

    ``(d - min) / (max - min)`` per row when ``max > min`` else zeros; same dtype as ``d``.
    """
    d_min = d.min(dim=1, keepdim=True).values
    d_max = d.max(dim=1, keepdim=True).values
    spread = d_max - d_min
    out = (d - d_min) / spread
    return torch.where(d_max > d_min, out, torch.zeros_like(out))


def kmeans_with_time_min_max(
    features: torch.Tensor,
    timestamp: torch.Tensor,
    cluster_num: int,
    alpha=2,
    max_iteration: int = 30,
    tol: float = 1e-4,
    *,
    init_indices: Optional[Sequence[int]] = None,
    exp_noise: Optional[torch.Tensor] = None,
    reseed_indices: Optional[Sequence[int]] = None,
    info: Optional[KMeansInfo] = None,
    exact: bool = False,
):
    """Temporal-semantic k-means, model/kmeans_with_time.py:4-137.

    Returns ``(features_kmeans [k,P,D] fp32, times_kmeans [k], assignments [T] int64)`` or,
    when ``T <= cluster_num``, ``(features_fp32 [T,P,D], timestamp[:k], None)`` (:30-32).

    Injection (keyword only, not in the reference):
      init_indices    k frame indices used as initial centres instead of k-means++ sampling.
      exp_noise       [k-1, T] fp32 Exp(1) draws; the n-th SAMPLING round picks
                      argmax(probs / exp_noise[n]), which is what ``torch.multinomial(probs, 1)``
                      computes internally (degenerate rounds, :55-57, consume no row).
      reseed_indices  frame indices consumed (in order) by empty-cluster re-seeding (:118).
      exact           evaluate the SAME algorithm with feature distances and centroid means in float64
                      (distances then rounded to the fp32 ``cdist`` output the reference works with): a
                      machine-independent cross-check that the fp32 sgemm rounding noise of the reference
                      (~1e-6 on the normalised distances) does not decide any assignment of a test case.
    """
    info = info if info is not None else KMeansInfo()
    features = features.to(dtype=torch.float32)  # :22
    if not isinstance(timestamp, torch.Tensor):  # :26-27
        timestamp = torch.tensor(timestamp, dtype=torch.float32)
    T, P, D = features.shape  # :29
    if T <= cluster_num:  # :30-32
        return features, timestamp[:cluster_num], None
    X = features.view(T, P * D)  # :35
    if exact:
        X = X.double()

    def _cdist(a, b):
        return torch.cdist(a, b, p=2).to(torch.float32) if exact else torch.cdist(a, b, p=2)

    # ---- k-means++ init on feature distance only (:41-62) -------------------------------
    if init_indices is not None:
        centers_indices = [int(i) for i in init_indices]
        assert len(centers_indices) == cluster_num
    else:
        centers_indices = [random.randint(0, T - 1)]  # :41
        n_sampled = 0  # multinomial calls so far == rows of exp_noise consumed
        while len(centers_indices) < cluster_num:  # :44
            dist_to_chosen = _cdist(X, X[centers_indices])  # :46-48
            nearest = dist_to_chosen.min(dim=1).values  # :50
            probs = nearest ** 2  # :53
            probs_sum = probs.sum()  # :54
            if probs_sum.item() == 0:  # :55-57
                info.degenerate_rounds += 1
                nxt = random.randint(0, T - 1)
            else:
                probs = probs / probs_sum  # :59
                if exp_noise is not None:
                    nxt = int(torch.argmax(probs / exp_noise[n_sampled]).item())
                else:
                    nxt = int(torch.multinomial(probs, 1).item())  # :60
                n_sampled += 1
            centers_indices.append(nxt)  # :62
    info.init_indices = list(centers_indices)

    centers_features = X[centers_indices]  # :65
    centers_times = timestamp[centers_indices]  # :66
    reseed_iter = iter(reseed_indices) if reseed_indices is not None else None

    assignments = None
    for iteration in range(max_iteration):  # :71
        d_feat = _cdist(X, centers_features)  # :73   fp32 [T,k]
        d_time = torch.abs(timestamp.unsqueeze(1) - centers_times.unsqueeze(0))  # :76
        norm_feat = _row_minmax_normalise(d_feat)  # :86-91
        norm_time = _row_minmax_normalise(d_time)  # :94-99
        # :101  fp32**2 + alpha*(time dtype)**2 promotes to the time dtype (float64 in
        # production), sqrt there, then rounds into the fp32 ``final_distances`` (:79).
        final = torch.sqrt(norm_feat ** 2 + alpha * (norm_time ** 2)).to(torch.float32)
        assignments = final.argmin(dim=1)  # :104 (first minimum)
        if cluster_num > 1:
            two = final.topk(2, dim=1, largest=False).values
            info.min_margin = min(info.min_margin, float((two[:, 1] - two[:, 0]).min()))

        new_feat = torch.zeros_like(centers_features)  # :107
        new_time = torch.zeros_like(centers_times)  # :108
        for i in range(cluster_num):  # :110-120
            member = assignments == i
            pts = X[member]
            if pts.size(0) > 0:
                new_feat[i] = pts.mean(dim=0)
                new_time[i] = timestamp[member].mean()
            else:
                if reseed_iter is not None:
                    r = int(next(reseed_iter))
                else:
                    r = random.randint(0, T - 1)  # :118
                info.reseed_indices.append(r)
                new_feat[i] = X[r]
                new_time[i] = timestamp[r]
        shift_feat = torch.norm(new_feat - centers_features, p=2, dim=1).sum()  # :123
        shift_time = torch.norm(new_time - centers_times, p=2).sum()  # :124
        shift = shift_feat + shift_time  # :125
        centers_features, centers_times = new_feat, new_time  # :127-128
        info.iterations = iteration + 1
        if shift <= tol:  # :130
            info.converged = True
            break
    return centers_features.to(torch.float32).view(cluster_num, P, D), centers_times, assignments  # :134-137


def select_additional_frames(
    cls_feature: torch.Tensor,
    long_memory: torch.Tensor,
    cluster_assignments: torch.Tensor,
    additional_frame_num: int,
) -> List[torch.Tensor]:
    """Representative frames per event, model/cogreasoner_chat.py:50-64.

    Events with <= n members keep all of them; otherwise the n members nearest (L2) to the
    event centroid, in ``topk(largest=False)`` order.
    """
    X = cls_feature.to(dtype=torch.float32)
    X = X.view(X.shape[0], -1)
    C = long_memory.view(long_memory.shape[0], -1)
    out = []
    for i in range(C.shape[0]):
        member = cluster_assignments == i
        idx = torch.nonzero(member, as_tuple=True)[0]
        if idx.numel() <= additional_frame_num:
            out.append(idx)
        else:
            dist = torch.cdist(X[member], C[i].unsqueeze(0)).squeeze(1)
            _, top = torch.topk(dist, k=additional_frame_num, largest=False)
            out.append(idx[top])
    return out


def event_gather(features: torch.Tensor, timestamps: torch.Tensor, assignments: torch.Tensor, k: int):
    """Chronological per-event gather, model/cogreasoner_chat.py:286-296.

    Returns ``(events, event_times)``: lists of ``[n_i*P, D]`` feature tensors (input dtype)
    and ``[n_i]`` timestamp tensors, frames in ascending index order within an event.
    """
    events, times = [], []
    for i in range(k):
        idx = torch.nonzero(assignments == i, as_tuple=True)[0]
        events.append(features[idx].reshape(-1, features.shape[-1]))
        times.append(timestamps[idx])
    return events, times


def cosine_low_events(question_embedding: torch.Tensor, events_repr: torch.Tensor, threshold: float = 0.45):
    """model/cogreasoner_chat.py:325-330 -- cosine in the embeddings' dtype (bf16 in
    production, every ATen op rounding to bf16), ``cos < threshold`` with the Python scalar
    rounded to that dtype.  Returns ``(cos [k], low_event_indices LongTensor)``.
    """
    cos = torch.nn.functional.cosine_similarity(question_embedding[0], events_repr, dim=1)
    return cos, torch.where(cos < threshold)[0]


def embed_scatter(input_ids: torch.Tensor, image_token_index: int, embed_weight: torch.Tensor, features: torch.Tensor):
    """model/cogreasoner_chat.py:303-305 (event-summary prompt) and :567-572 (final prompt) --
    ``embed_tokens(input_ids).clone()``, then the rows of the <image> ids become ``embeds * 0.0 + features`` (two ATen ops,
    each rounding in the embedding dtype).  ``input_ids`` 1-D."""
    inputs_embeds = torch.nn.functional.embedding(input_ids, embed_weight).clone()
    image_selected = input_ids == image_token_index
    inputs_embeds[image_selected] = inputs_embeds[image_selected] * 0.0 + features.to(inputs_embeds.dtype)
    return inputs_embeds


def event_pool(hidden_states: Sequence[torch.Tensor]) -> torch.Tensor:
    """model/cogreasoner_chat.py:317-318 (and :322 for the question) -- ``torch.mean(last_hidden_state, dim=1)`` of every
    ``[1, L_i, Dl]`` output, concatenated to ``[k, Dl]``."""
    return torch.cat([torch.mean(h, dim=1) for h in hidden_states], dim=0)


def minor_frames(assignments: torch.Tensor, low_events: torch.Tensor, representatives: List[torch.Tensor]) -> List[int]:
    """model/cogreasoner_chat.py:285, :331-333 -- frames of low-relevance events that are not
    representatives, ascending."""
    keep = set(torch.cat(representatives, dim=0).view(-1).tolist()) if len(representatives) else set()
    low = set(low_events.tolist())
    return [i for i, a in enumerate(assignments.tolist()) if a in low and i not in keep]


def compress_unimportant_events(mm_features: torch.Tensor, patch_num: int, minor_frame_indices: Sequence[int]):
    """model/cogreasoner_chat.py:434-447 -- clone, then token 0 of every minor frame becomes
    the mean over that frame's tokens (computed in the tensor's dtype: fp32 accumulate, one
    rounding)."""
    total, dim = mm_features.shape
    if total % patch_num != 0:
        raise ValueError(f"total patches ({total}) not divisible by patch_num ({patch_num})")
    out = mm_features.view(total // patch_num, patch_num, dim).clone()
    for f in minor_frame_indices:
        out[f, 0, :] = out[f].mean(dim=0)
    return out.view(-1, dim)


def get_compression_mask(
    pixel_values: torch.Tensor,
    batched_num_patches: torch.Tensor,
    grid_sizes: torch.Tensor,
    merge_sizes: torch.Tensor,
    modals: List[str],
    threshold: float = 0.1,
    min_tokens: int = 1,
    minor_frame_indices: Optional[Sequence[int]] = None,
) -> torch.Tensor:
    """model/cogreasoner_chat.py:383-432.

    video item (t,h,w), t>1: view pixels as [t, P, m*m*C]; token kept iff
    mean|x_f - x_{f-1}| * 255 > threshold (dtype arithmetic; first frame always kept; frames
    with < min_tokens kept get their first min_tokens); minor frames keep only token 0.
    image / single-frame video: all kept.  anything else: contributes no tokens.
    """
    minor = set(int(i) for i in minor_frame_indices) if minor_frame_indices is not None else None
    chunks = pixel_values.split(grid_sizes.prod(dim=1).tolist(), dim=0)
    masks, frame0 = [], 0
    for images, n_patches, grid, merge, modal in zip(chunks, batched_num_patches, grid_sizes, merge_sizes, modals):
        t, h, w = (int(v) for v in grid)
        merge = int(merge)
        if modal == "image" or (modal == "video" and t == 1):
            masks.append(torch.ones((int(n_patches),), dtype=torch.bool))
        elif modal == "video":
            if minor is None:
                raise TypeError("argument of type 'NoneType' is not iterable")  # :419 behaviour
            img = images.view(t, (h // merge) * (w // merge), -1)
            diff = torch.abs(img[1:] - img[:-1]).mean(dim=-1) * 255
            diff = torch.cat([torch.full_like(diff[0:1], threshold + 1), diff], dim=0)
            mask = diff > threshold
            short = torch.nonzero(mask.sum(dim=1) < min_tokens)[:, 0]
            mask[short, :min_tokens] = 1
            for f in range(t):
                if frame0 + f in minor:
                    mask[f, 0] = True
                    mask[f, 1:] = False
            masks.append(mask.flatten())
        else:
            masks.append(torch.ones((0,), dtype=torch.bool))
        frame0 += t
    return torch.cat(masks)


def compress_visual_tokens(
    compression_mask: torch.Tensor,
    mm_features: torch.Tensor,
    input_ids: torch.Tensor,
    image_token_index: int,
    attention_mask: Optional[torch.Tensor] = None,
    position_ids: Optional[torch.Tensor] = None,
    labels: Optional[torch.Tensor] = None,
):
    """model/cogreasoner_chat.py:449-476 -- drop the masked-out visual tokens and the
    ``<image>`` placeholders at the same positions; order preserved."""
    mm_features = mm_features[compression_mask]
    image_selected = input_ids == image_token_index
    keep = torch.logical_not(image_selected)
    keep[image_selected] = compression_mask
    input_ids = input_ids[keep]
    if attention_mask is not None:
        attention_mask = attention_mask[keep]
    if labels is not None:
        labels = labels[keep]
    if position_ids is not None:
        position_ids = position_ids[keep]
        starts = [0] + torch.nonzero(position_ids == 0)[:, 0].tolist()
        ends = starts[1:] + [len(input_ids)]
        position_ids = torch.cat([torch.arange(e - s) for s, e in zip(starts, ends)])
    return mm_features, input_ids, attention_mask, position_ids, labels


def encoder_tail(hidden_states: torch.Tensor, grid_sizes: torch.Tensor, merge_sizes: torch.Tensor, ln_weight: torch.Tensor,
                 ln_bias: torch.Tensor, eps: float) -> torch.Tensor:
    """model/modeling_videollama3_encoder.py:479-504 -- ``post_layernorm`` followed, per item, by the view / permute / reshape /
    permute / bilinear ``interpolate`` / permute / view chain that merges every ``merge x merge`` patch block into one token.
    Runs on whatever device the inputs live on (the parity tests run it with the CUDA ATen kernels the reference itself uses)."""
    h = torch.nn.functional.layer_norm(hidden_states, (hidden_states.shape[-1],), ln_weight, ln_bias, eps)  # :479
    chunks = h.split(grid_sizes.prod(dim=1).tolist(), dim=0)  # :481
    outs = []
    for hs, grid, m in zip(chunks, grid_sizes.tolist(), merge_sizes.tolist()):  # :484-502
        t, gh, gw = grid
        c = hs.shape[-1]
        hs = hs.view(t, gh // m, gw // m, m, m, c).permute(0, 1, 3, 2, 4, 5)
        hs = hs.reshape(t, gh, gw, c).permute(0, 3, 1, 2)
        hs = torch.nn.functional.interpolate(hs, size=(gh // m, gw // m), mode="bilinear")
        outs.append(hs.permute(0, 2, 3, 1).reshape(-1, c))
    return torch.cat(outs, dim=0)  # :504


# --------------------------------------------------------------------------------------
# Whole-path drivers (what bench.py's CPU arm times).  The frozen-LLM forwards of
# cogreasoner_chat.py:302-323 are out of scope: their outputs (event / question
# embeddings) are inputs here.
# --------------------------------------------------------------------------------------
@dataclass
class Phase1Result:
    k: int
    centroids: Optional[torch.Tensor]
    centroid_times: Optional[torch.Tensor]
    assignments: Optional[torch.Tensor]
    representatives: List[torch.Tensor]
    events: List[torch.Tensor]
    event_times: List[torch.Tensor]
    info: KMeansInfo


@dataclass
class Phase2Result:
    cos: torch.Tensor
    low_events: torch.Tensor
    minor: List[int]
    mask: torch.Tensor
    features: torch.Tensor
    input_ids: torch.Tensor
    attention_mask: Optional[torch.Tensor]


def phase1(
    mm_features: torch.Tensor,
    total_image_num: int,
    timestamps: torch.Tensor,
    *,
    cluster_num: Optional[int] = None,
    n_representatives: int = 2,
    min_clusters: int = 9,
    **kmeans_kw,
) -> Optional[Phase1Result]:
    """cogreasoner_chat.py:279-301: k = ceil(T/15), gate k > 9, k-means, representatives,
    per-event chronological gather.  ``None`` when the gate closes (:281-282)."""
    T = total_image_num
    feats = mm_features.view(T, mm_features.shape[0] // T, mm_features.shape[1])  # :279
    k = math.ceil(T / 15) if cluster_num is None else cluster_num  # :280
    if cluster_num is None and k <= min_clusters:  # :281
        return None
    info = KMeansInfo()
    cent, ct, assign = kmeans_with_time_min_max(feats, timestamps, k, info=info, **kmeans_kw)  # :283
    reps = select_additional_frames(feats, cent, assign, n_representatives)  # :284
    events, etimes = event_gather(feats, timestamps, assign, k)  # :286-296
    return Phase1Result(k, cent, ct, assign, reps, events, etimes, info)


def phase2(
    mm_features: torch.Tensor,
    p1: Optional[Phase1Result],
    events_repr: Optional[torch.Tensor],
    question_embedding: Optional[torch.Tensor],
    pixel_values: torch.Tensor,
    grid_sizes: torch.Tensor,
    merge_sizes: torch.Tensor,
    modals: List[str],
    input_ids: torch.Tensor,
    image_token_index: int,
    attention_mask: Optional[torch.Tensor] = None,
    *,
    threshold: float = 0.45,
    pixel_threshold: float = 0.1,
    min_tokens: int = 1,
) -> Phase2Result:
    """cogreasoner_chat.py:325-333 then :551-564."""
    if p1 is None:
        cos, low, minor = torch.empty(0), torch.empty(0, dtype=torch.long), []
        T = int(grid_sizes[:, 0].sum())
    else:
        cos, low = cosine_low_events(question_embedding, events_repr, threshold)
        assert cos.shape[0] == p1.k  # :327
        minor = minor_frames(p1.assignments, low, p1.representatives)
        T = p1.assignments.shape[0]
    patch_num = mm_features.shape[0] // T
    feats = compress_unimportant_events(mm_features, patch_num, minor)  # :551
    n_patches = grid_sizes.prod(dim=1).div(merge_sizes ** 2).long()  # :545
    mask = get_compression_mask(
        pixel_values, n_patches, grid_sizes, merge_sizes, modals,
        threshold=pixel_threshold, min_tokens=min_tokens, minor_frame_indices=minor,
    )  # :552
    feats, ids, am, _, _ = compress_visual_tokens(mask, feats, input_ids, image_token_index, attention_mask)  # :562
    return Phase2Result(cos, low, minor, mask, feats, ids, am)
