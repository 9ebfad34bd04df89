"""One var-len forward of the frozen LLM over the k packed event-summary prompts.

The reference runs the k per-event summary prompts through the frozen Qwen2 decoder one by one, batch 1 each
(``model/cogreasoner_chat.py:306-316``).  The drop-in ``select_events_based_on_summary`` already embeds all k prompts into ONE
packed ``[sum(len), D]`` buffer with ``cu_seqlens`` (``csb_embed_scatter_ordered``); this adapter drives the *untouched* HF
decoder modules (their weights, norms, rotary embedding, projections, MLPs -- nothing is re-implemented or re-trained) over
that packed buffer with a var-len causal attention call, and the packed last hidden states go straight into
``csb_relevance_select`` (mean pool + cosine + selection in one launch):

    model.csb_batched_summary_forward = make_varlen_summary_forward(model.get_model())

What is ours here is only the orchestration: per-token position ids that restart at every sequence (what ``position_ids=None``
gives each batch-1 call of the reference), the layer loop, and the choice of the var-len attention kernel, which is LIBRARY code:
``flash_attn.flash_attn_varlen_func`` when importable, else ``torch.nn.attention.varlen.varlen_attn``.  ``attention="sdpa_loop"``
(per-sequence ``scaled_dot_product_attention``) exists to validate the orchestration against ``Qwen2Model.forward`` itself.
The packed forward is mathematically the reference's k forwards; in bf16 its hidden states differ from the sequential calls by
kernel-level rounding (different attention kernels), so it is an opt-in, not the default of the drop-in.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

__all__ = ["make_varlen_summary_forward", "packed_position_ids"]


def packed_position_ids(cu_seqlens: torch.Tensor, n: int) -> torch.Tensor:
    """[n] int64: 0 .. len_j - 1 inside every sequence j of the packed batch."""
    cu = cu_seqlens.to(torch.int64)
    lens = cu[1:] - cu[:-1]
    starts = torch.repeat_interleave(cu[:-1], lens, output_size=n)
    return torch.arange(n, device=cu_seqlens.device, dtype=torch.int64) - starts


def _rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def _pick_attention(name: Optional[str]) -> Callable:
    """-> fn(q [N,H,hd], k [N,Hkv,hd], v [N,Hkv,hd], cu [n+1] int32, max_len, scale) -> [N,H,hd], causal inside each sequence."""
    if name in (None, "flash_attn"):
        try:
            from flash_attn import flash_attn_varlen_func

            def fa(q, k, v, cu, max_len, scale):
                return flash_attn_varlen_func(q, k, v, cu, cu, max_len, max_len, dropout_p=0.0, softmax_scale=scale, causal=True)
            return fa
        except Exception:
            if name == "flash_attn":
                raise
    if name in (None, "torch_varlen"):
        try:
            from torch.nn.attention.varlen import varlen_attn

            def tv(q, k, v, cu, max_len, scale):
                if k.shape[1] != q.shape[1]:  # grouped-query attention: expand the kv heads
                    rep = q.shape[1] // k.shape[1]
                    k, v = k.repeat_interleave(rep, dim=1), v.repeat_interleave(rep, dim=1)
                return varlen_attn(q, k, v, cu, cu, max_len, max_len, scale=scale, window_size=(-1, 0))
            return tv
        except Exception:
            if name == "torch_varlen":
                raise
    if name == "sdpa_loop":
        def loop(q, k, v, cu, max_len, scale):
            out = torch.empty_like(q)
            rep = q.shape[1] // k.shape[1]
            b = cu.tolist()
            for j in range(len(b) - 1):
                s, e = b[j], b[j + 1]
                if e == s:
                    continue
                qj = q[s:e].transpose(0, 1).unsqueeze(0)
                kj = k[s:e].repeat_interleave(rep, dim=1).transpose(0, 1).unsqueeze(0)
                vj = v[s:e].repeat_interleave(rep, dim=1).transpose(0, 1).unsqueeze(0)
                oj = torch.nn.functional.scaled_dot_product_attention(qj, kj, vj, is_causal=True, scale=scale)
                out[s:e] = oj[0].transpose(0, 1)
            return out
        return loop
    raise RuntimeError("no var-len attention kernel available (flash_attn / torch.nn.attention.varlen); pass attention='sdpa_loop' "
                       "for the per-sequence loop")


def make_varlen_summary_forward(llm, *, attention: Optional[str] = None) -> Callable[[torch.Tensor, torch.Tensor], torch.Tensor]:
    """Build ``csb_batched_summary_forward`` for an HF Qwen2-style decoder stack (``llm`` = what ``self.get_model()`` returns:
    ``.layers[i].{input_layernorm, self_attn.{q,k,v,o}_proj, post_attention_layernorm, mlp}``, ``.rotary_emb``, ``.norm``).

    The returned callable maps ``(inputs_embeds [N, D], cu_seqlens [k+1] int32) -> last hidden states [N, Dl]`` where the N rows
    are the k prompts back to back; sequence j attends causally inside ``[cu[j], cu[j+1])`` only."""
    for name in ("layers", "norm", "rotary_emb"):
        if not hasattr(llm, name):
            raise TypeError(f"make_varlen_summary_forward: the model has no .{name} (expected an HF Qwen2-style decoder stack)")
    attn_fn = _pick_attention(attention)
    att0 = llm.layers[0].self_attn
    if getattr(att0, "sliding_window", None):
        raise NotImplementedError("make_varlen_summary_forward: sliding-window attention layers are not supported")

    @torch.no_grad()
    def forward(inputs_embeds: torch.Tensor, cu_seqlens: torch.Tensor) -> torch.Tensor:
        if inputs_embeds.dim() != 2:
            raise ValueError("inputs_embeds must be the packed [sum(len), D] buffer")
        n = inputs_embeds.shape[0]
        cu = cu_seqlens.to(device=inputs_embeds.device, dtype=torch.int32)
        pos = packed_position_ids(cu, n)
        max_len = int((cu[1:] - cu[:-1]).max().item())
        h = inputs_embeds
        cos, sin = llm.rotary_emb(h.unsqueeze(0), pos.unsqueeze(0))  # [1, N, hd] each, in h's dtype
        cos, sin = cos[0].unsqueeze(1), sin[0].unsqueeze(1)  # [N, 1, hd]: broadcast over heads
        for layer in llm.layers:
            att = layer.self_attn
            hd = att.head_dim
            x = layer.input_layernorm(h)
            q = att.q_proj(x).view(n, -1, hd)
            k = att.k_proj(x).view(n, -1, hd)
            v = att.v_proj(x).view(n, -1, hd)
            q = q * cos + _rotate_half(q) * sin
            k = k * cos + _rotate_half(k) * sin
            o = attn_fn(q.contiguous(), k.contiguous(), v.contiguous(), cu, max_len, float(att.scaling))
            h = h + att.o_proj(o.reshape(n, -1))
            h = h + layer.mlp(layer.post_attention_layernorm(h))
        return llm.norm(h)

    return forward
