"""Encoder-tail fusion: the step immediately upstream of the compression path.

``Videollama3VisionEncoderModel.forward`` (model/modeling_videollama3_encoder.py:477-506) ends with ``post_layernorm`` and, per
visual item, a view / permute / reshape / permute / ``F.interpolate(mode='bilinear')`` / permute / view chain that merges every
``merge x merge`` patch block into one token; ``encode_images`` (model/cogreasoner_chat.py:264-276) then applies the projector.
The chain touches the [n_patches, C] tensor five times; ``csb_ln_merge`` does LayerNorm + merge in ONE pass (the patches of a
block are consecutive rows already), reading the hidden states once and writing the 4x smaller merged tensor once.

    encoder_tail(vision_encoder, hidden_states, grid_sizes, merge_sizes)  -> what lines :479-504 return
    FusedEncoderTailMixin.forward                                         -> drop-in for Videollama3VisionEncoderModel.forward
    install_encoder_tail(encoder_module)                                  -> rebinds it on the imported reference module

The frozen transformer layers, the embeddings and the projector (two library GEMMs + GELU on the frozen weights, 3.7 TFLOP per
config-2 stream: tensor-bound library work, not part of this path) are called as they are.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["encoder_tail", "FusedEncoderTailMixin", "install_encoder_tail"]


def encoder_tail(post_layernorm: torch.nn.LayerNorm, hidden_states: torch.Tensor, grid_sizes: torch.Tensor, merge_sizes) -> torch.Tensor:
    """:479-504 -- ``post_layernorm`` then the per-item bilinear merge, concatenated.  ``hidden_states`` [sum(t*h*w), C] bf16 on
    the device; items that share a merge size are processed in one launch (the op is item-independent)."""
    if not hidden_states.is_cuda:
        raise RuntimeError("cogstream_b200: CPU tensor given; there is no CPU fallback")
    if hidden_states.dtype != torch.bfloat16:
        raise TypeError(f"cogstream_b200.encoder_tail: hidden states must be bf16 (the tower runs in bf16), got {hidden_states.dtype}")
    w, b = post_layernorm.weight, post_layernorm.bias
    if w is None or b is None or w.dtype != torch.bfloat16:
        raise TypeError("cogstream_b200.encoder_tail: post_layernorm must have bf16 affine parameters")
    h = hidden_states.contiguous()
    counts = grid_sizes.prod(dim=1).tolist()  # the reference syncs here too (:481)
    merges = [int(m) for m in (merge_sizes.tolist() if isinstance(merge_sizes, torch.Tensor) else merge_sizes)]
    for (t, gh, gw), m in zip(grid_sizes.tolist(), merges):
        if gh % m or gw % m:
            raise RuntimeError(f"shape '[{t}, {gh // m}, {gw // m}, {m}, {m}, -1]' is invalid for input of size {t * gh * gw}")  # :486 view()
    outs, row0, i = [], 0, 0
    while i < len(counts):  # runs of items with the same merge size -> one launch each
        j, rows = i, 0
        while j < len(counts) and merges[j] == merges[i]:
            rows += counts[j]
            j += 1
        outs.append(ops.ln_merge(h[row0: row0 + rows], merges[i], w.detach(), b.detach(), post_layernorm.eps))
        row0 += rows
        i = j
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)


class FusedEncoderTailMixin:
    """``forward`` of ``Videollama3VisionEncoderModel`` (:477-506) with the tail fused; same signature and return."""

    def forward(self, pixel_values, grid_sizes, merge_sizes=None) -> torch.Tensor:
        hidden_states = self.embeddings(pixel_values)
        hidden_states = self.encoder(hidden_states, grid_sizes, merge_sizes)
        return encoder_tail(self.post_layernorm, hidden_states, grid_sizes, merge_sizes)


def install_encoder_tail(encoder_module, *, class_name: str = "Videollama3VisionEncoderModel"):
    """Rebind ``forward`` of the imported reference ``model.modeling_videollama3_encoder`` module's encoder class."""
    cls = getattr(encoder_module, class_name)
    cls.forward = FusedEncoderTailMixin.forward
    return cls
