"""Drop-ins for the compression functions/methods of the reference's ``model/cogreasoner_chat.py``.

Same names, argument meaning, return types and error behaviour as the reference:

    select_additional_frames(cls_feature, long_memory, cluster_assignments, additional_frame_num)   :50-64
    Videollama3MetaForCausalLM.select_events_based_on_summary(self, mm_features, total_image_num, timestamps)  :278-333
    Videollama3MetaForCausalLM.compress_unimportant_events(self, mm_features, patch_num, minor_frame_indices)  :434-447
    Videollama3MetaForCausalLM._get_compression_mask(self, pixel_values, batched_num_patches, grid_sizes,
                                                     merge_sizes, modals, threshold=0.1, min_tokens=1,
                                                     minor_frame_indices=None)                       :383-432
    Videollama3MetaForCausalLM._compress_visual_tokens(self, compression_mask, mm_features, input_ids,
                                                     attention_mask=None, position_ids=None, labels=None) :449-476

``VisualStreamCompressionMixin`` carries the four methods; ``install(chat_module)`` rebinds them (and the
two module-level functions) on an imported reference module, which is all a maintainer has to do
(INTEGRATION.md).  The frozen LLM forwards inside ``select_events_based_on_summary`` (:302-323) are
called exactly like the reference calls them -- the pretrained model is untouched.
Inference only: the CUDA ops have no backward (``forward_train`` keeps the reference path).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch

from . import ops
from .kmeans_with_time import kmeans_with_time_min_max

__all__ = ["select_additional_frames", "create_visual_summary_prompt", "VisualStreamCompressionMixin", "install"]


def _frames_bf16(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("cogstream_b200: CPU tensor given; there is no CPU fallback")
    if x.dtype != torch.bfloat16:
        y = x.to(torch.bfloat16)
        if not torch.equal(y.to(x.dtype), x):
            raise NotImplementedError("cogstream_b200: features must be bf16 (or exactly bf16-representable)")
        x = y
    return x.contiguous()


def select_additional_frames(cls_feature, long_memory, cluster_assignments, additional_frame_num):
    """Per event the ``additional_frame_num`` member frames nearest (L2) to the event centroid, nearest first;
    events with no more members than that keep all of them (ascending).  cogreasoner_chat.py:50-64.

    One streaming pass over the features computes every frame's distance to its own centroid
    (``csb_dist_to_assigned``), a warp per event then selects.  Ties pick the lowest frame index
    (``torch.topk``'s tie order is implementation-defined).
    """
    x = _frames_bf16(cls_feature)
    T = x.shape[0]
    k = long_memory.shape[0]
    cent = long_memory.reshape(k, -1).to(torch.float32).contiguous()
    assign = cluster_assignments.to(torch.int64).contiguous()
    dist = ops.dist_to_assigned(x.view(T, -1), cent, assign)
    reps, counts, _ = ops.select_representatives(dist, assign, k, int(additional_frame_num))
    counts_h = counts.tolist()
    return [reps[i, : counts_h[i]] for i in range(k)]


def _mean_pool(hidden: torch.Tensor, cu_seqlens: torch.Tensor) -> torch.Tensor:
    """``torch.mean(last_hidden_state, dim=1)`` (:317, :322) of the sequences of a packed ``[sum(len), Dl]`` tensor."""
    if not hidden.is_cuda:
        raise RuntimeError("cogstream_b200: the LLM's hidden states are on the CPU; there is no CPU fallback")
    if hidden.dtype not in (torch.bfloat16, torch.float32) or hidden.shape[1] % 8 != 0:
        raise TypeError(f"cogstream_b200: hidden states must be bf16 / fp32 with a width that is a multiple of 8, got "
                        f"{hidden.dtype} x {hidden.shape[1]}")
    return ops.seq_mean_pool(hidden.contiguous(), cu_seqlens)


def _embed_weight(llm, D: int) -> torch.Tensor:
    """The frozen LLM's embedding table as a plain tensor (the fused embed + <image>-row replacement reads it directly)."""
    w = getattr(getattr(llm, "embed_tokens", None), "weight", None)
    if not (isinstance(w, torch.Tensor) and w.is_cuda and w.dtype in (torch.bfloat16, torch.float32) and w.is_contiguous()):
        raise TypeError("cogstream_b200: get_model().embed_tokens must be an nn.Embedding with a contiguous bf16 / fp32 CUDA weight "
                        "(wrapped or sharded embedding modules are not supported; there is no eager fallback)")
    if D is not None and w.shape[1] != D:
        raise ValueError(f"cogstream_b200: feature width {D} != embedding width {w.shape[1]}")
    return w


def create_visual_summary_prompt(P, timestamps, image_token="<image>"):
    """Prompt of the per-event summary forward (text layout of cogreasoner_chat.py:93-119; it has to match
    token for token, since the frozen LLM's pooled hidden state is the event embedding)."""
    n = len(timestamps)
    per_frame = image_token * (P // n)
    frames = []
    for t in range(n):
        v = timestamps[t].item() if isinstance(timestamps[t], torch.Tensor) else float(timestamps[t])
        frames.append(f"Time {v:.1f}s:{per_frame}" + ("," if t < n - 1 else ""))
    instruction = (
        "Concisely list the key points of the event shown in the timestamped images, adhering strictly and honestly to "
        "the visual content. For each key point, identify relevant objects or actions, note any visible text, and specify "
        "the approximate timestamp(s). Provide an overview focusing on these key timestamped points."
    )
    return (
        "<|im_start|>system\n"
        "You are a helpful assistant specializing in summarizing events from timestamped visual data.<|im_end|>\n"
        "<|im_start|>user\n" + "".join(frames) + "\n" + instruction + "<|im_end|>\n" + "<|im_start|>assistant"
    )


class VisualStreamCompressionMixin:
    """The four compression methods of ``Videollama3MetaForCausalLM`` on the B200 kernels."""

    # hard-coded in the reference; exposed here with the reference values as defaults
    csb_frames_per_event = 15  # :280
    csb_min_events = 9  # :281
    csb_n_representatives = 2  # :284
    csb_relevance_threshold = 0.45  # :329

    # ------------------------------------------------------------------------------------------
    def select_events_based_on_summary(self, mm_features, total_image_num, timestamps):
        feats = _frames_bf16(mm_features)
        T = int(total_image_num)
        P = feats.shape[0] // T
        D = feats.shape[1]
        features = feats.view(T, P, D)  # :279
        k = math.ceil(T / self.csb_frames_per_event)  # :280
        if k <= self.csb_min_events:  # :281-282
            return []
        dev = feats.device
        ts = timestamps if isinstance(timestamps, torch.Tensor) else torch.tensor(timestamps)
        ts = ts.to(dev)
        # csb_kmeans_kwargs: optional keyword-only extras of the drop-in k-means (init / reseed injection for tests)
        _, _, assign, info = kmeans_with_time_min_max(features, ts, k, return_info=True, return_centroids=False,
                                                      **getattr(self, "csb_kmeans_kwargs", {}))  # :283
        order, ev_off = ops.event_order(assign, k)
        # chronological per-event gather (:286-296): NOT materialised; the embedding pass below reads the frame tokens through
        # `order` (event-major, ascending frame index inside an event)
        off_h = ev_off.tolist()
        ts_sorted = ts[order.long()].tolist()
        # prompts of all events first (:294-301), packed back to back (cu_seqlens layout)
        ids_list, am_list = [], []
        for i in range(k):
            lo, hi = off_h[i], off_h[i + 1]
            prompt = create_visual_summary_prompt((hi - lo) * P, ts_sorted[lo:hi])
            inputs = self.tokenizer(prompt, return_tensors="pt")
            ids_list.append(inputs["input_ids"].view(-1))
            am_list.append(inputs["attention_mask"])
        lens = [int(t.numel()) for t in ids_list]
        cu_h = [0]
        for n_i in lens:
            cu_h.append(cu_h[-1] + n_i)
        ids_all = torch.cat(ids_list).to(dev)
        llm = self.get_model()
        # :303-305 for all k prompts in ONE pass: the j-th <image> id overall takes token j of the event-major frame order
        embeds_all, _ = ops.embed_scatter(ids_all, self.config.image_token_index, _embed_weight(llm, D), feats, frame_order=order,
                                          tokens_per_frame=P)
        cu_all = torch.tensor(cu_h, dtype=torch.int32).to(dev)
        hidden_all = None
        events_repr = None
        batched = getattr(self, "csb_batched_summary_forward", None)
        if callable(batched):
            # one var-len forward of the frozen LLM over the packed prompts (cogstream_b200.summary_forward
            # builds one for HF Qwen2-style models; any callable (inputs_embeds [sum(len), D], cu_seqlens) -> [sum(len), Dl] works)
            hidden_all = batched(embeds_all, cu_all)
            if not hidden_all.is_cuda:
                raise RuntimeError("cogstream_b200: csb_batched_summary_forward returned a CPU tensor; there is no CPU fallback")
            hidden_all = hidden_all.contiguous()
        else:
            cu_pairs = torch.tensor([[0, n_i] for n_i in lens] + [[0, 0]], dtype=torch.int32).to(dev)  # per-call [0, L_i] rows
            events_repr = []
            for i in range(k):  # :306-317, the frozen LLM is called exactly like the reference does
                e_i = embeds_all[cu_h[i]: cu_h[i + 1]]
                out = llm(position_ids=None, attention_mask=am_list[i].to(dev).view(1, -1),
                          inputs_embeds=e_i.reshape(1, -1, e_i.shape[-1]))
                events_repr.append(_mean_pool(out.last_hidden_state[0], cu_pairs[i]))  # :317
            events_repr = torch.cat(events_repr, dim=0)  # :318
        inputs = self.tokenizer(self.current_question, padding=True, truncation=True, return_tensors="pt", max_length=128)
        model_device = next(llm.parameters()).device
        inputs = {kk: v.to(model_device) for kk, v in inputs.items()}
        q_hidden = llm(**inputs).last_hidden_state  # :319-322
        if q_hidden.shape[0] == 1:
            q_cu = torch.tensor([0, q_hidden.shape[1]], dtype=torch.int32).to(q_hidden.device)
            question = _mean_pool(q_hidden[0], q_cu)
        else:
            question = torch.mean(q_hidden, dim=1)

        # :284 representatives + :325-333 relevance + minor list: one launch (the event mean pool :317-318 fused in when the hidden
        # states came back packed)
        src = hidden_all if hidden_all is not None else events_repr.to(dev).contiguous()
        if src.dtype not in (torch.bfloat16, torch.float32):
            src = src.float()
        q_c = question[0].to(device=dev, dtype=src.dtype).contiguous()
        ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        if T <= 256 * k:
            if hidden_all is not None:
                b = ops.relevance_select(q_c, None, info.gram, ev_off, order, assign, ticket, k=k, n_rep=self.csb_n_representatives,
                                         threshold=self.csb_relevance_threshold, hidden=src, cu_seqlens=cu_all)
            else:
                b = ops.relevance_select(q_c, src, info.gram, ev_off, order, assign, ticket, k=k, n_rep=self.csb_n_representatives,
                                         threshold=self.csb_relevance_threshold)
            assert b["low"].shape[0] == k  # :327
            minor_list, n_minor = b["minor_list"], b["n_minor"]
        else:  # events of many hundreds of frames: the two-pass distance kernels
            ev_c = src if hidden_all is None else _mean_pool(src, cu_all)
            dist = ops.dist_to_assigned_gram(info.gram, assign, ev_off, order, k)
            _, _, rep_flags = ops.select_representatives(dist, assign, k, self.csb_n_representatives)
            _, low = ops.cosine_flags(q_c, ev_c, self.csb_relevance_threshold)
            assert low.shape[0] == k  # :327
            _, minor_list, n_minor = ops.minor_frames(assign, low, rep_flags)
        return minor_list[: int(n_minor.item())].tolist()  # :333

    # ------------------------------------------------------------------------------------------
    def compress_unimportant_events(self, mm_features, patch_num, minor_frame_indices):
        total_patches, embedding_dim = mm_features.shape
        if total_patches % patch_num != 0:  # :440-441
            raise ValueError(f"total patches ({total_patches}) not divisible by patches per frame ({patch_num})")
        out = _frames_bf16(mm_features).clone()  # :443
        if len(minor_frame_indices):
            frames = torch.tensor(list(minor_frame_indices), dtype=torch.int32, device=out.device)
            ops.pool_frames_inplace(out, int(patch_num), frames)  # :444-446
        return out if out.dtype == mm_features.dtype else out.to(mm_features.dtype)

    # ------------------------------------------------------------------------------------------
    def _get_compression_mask(self, pixel_values, batched_num_patches, grid_sizes, merge_sizes, modals,
                              threshold: float = 0.1, min_tokens: int = 1, minor_frame_indices: Optional[List[int]] = None):
        if not pixel_values.is_cuda:
            raise RuntimeError("cogstream_b200: CPU tensor given; there is no CPU fallback")
        dev = pixel_values.device
        pix = pixel_values if pixel_values.dtype in (torch.bfloat16, torch.float32) else pixel_values.float()
        pix = pix.contiguous()
        grids = grid_sizes.tolist()  # the reference syncs here too (:394)
        merges = merge_sizes.tolist()
        n_patches = [int(v) for v in batched_num_patches.tolist()]
        total_frames = sum(int(g[0]) for g in grids)
        minor_flags = None
        if minor_frame_indices is not None:
            minor_flags = torch.zeros(max(total_frames, 1), dtype=torch.uint8, device=dev)
            if len(minor_frame_indices):
                idx = torch.tensor([i for i in minor_frame_indices if 0 <= i < total_frames], dtype=torch.int64, device=dev)
                minor_flags[idx] = 1
        C = pix.shape[1]
        masks, row0, frame0 = [], 0, 0
        for (t, h, w), m, npat, modal in zip(grids, merges, n_patches, modals):
            rows = t * h * w
            if modal == "image" or (modal == "video" and t == 1):  # :401-403
                masks.append(torch.ones((npat,), dtype=torch.bool, device=dev))
            elif modal == "video":  # :405-425
                if minor_frame_indices is None:
                    raise TypeError("argument of type 'NoneType' is not iterable")  # what :419 raises in the reference
                P = (h // m) * (w // m)
                W = (rows * C) // (t * P)
                item = pix[row0: row0 + rows]
                mk = ops.pixel_mask(item, t, P, W, threshold=threshold, min_tokens=min_tokens, minor_flags=minor_flags,
                                    frame0=frame0)
                masks.append(mk.view(torch.bool))
            else:  # :427-430 pseudo image
                masks.append(torch.ones((0,), dtype=torch.bool, device=dev))
            row0 += rows
            frame0 += t
        return torch.cat(masks)  # :432

    # ------------------------------------------------------------------------------------------
    def _compress_visual_tokens(self, compression_mask, mm_features, input_ids, attention_mask=None, position_ids=None,
                                labels=None):
        feats = _frames_bf16(mm_features)
        mask_u8 = compression_mask.contiguous().view(torch.uint8) if compression_mask.dtype == torch.bool else \
            compression_mask.to(torch.uint8).contiguous()
        N = feats.shape[0]
        out, _, n_kept = ops.compact_tokens(feats, N, 1, mask_u8)  # :458
        ids = input_ids.contiguous()
        o_ids, o_am, o_lb, o_ps, n_out = ops.filter_ids(ids, self.config.image_token_index, mask_u8, attention_mask, labels,
                                                        position_ids)  # :459-474
        nk = int(n_kept.item())
        no, n_img = (int(v) for v in n_out.tolist())
        if n_img != mask_u8.numel():  # `text_masks[image_selected] = compression_mask` (:463) raises in the reference
            raise RuntimeError(f"shape mismatch: value tensor of shape [{mask_u8.numel()}] cannot be broadcast to indexing result of "
                               f"shape [{n_img}]")
        mm = out[:nk]
        if mm.dtype != mm_features.dtype:
            mm = mm.to(mm_features.dtype)

        def _cut(t, like):
            if t is None:
                return None
            t = t[:no]
            return t if like is None or t.dtype == like.dtype else t.to(like.dtype)

        return mm, o_ids[:no], _cut(o_am, attention_mask), _cut(o_ps, position_ids), _cut(o_lb, labels)


    # ------------------------------------------------------------------------------------------
    def prepare_inputs_labels_for_multimodal(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                                             labels=None, pixel_values=None, grid_sizes=None, merge_sizes=None, modals=None,
                                             total_image_num=0, if_visual=True):
        """The call site (cogreasoner_chat.py:509-585) with the same signature and return tuple.  The vision encoder,
        ``_get_valid_visual_tokens`` and ``_maybe_truncate_visual_tokens`` are the host class's own; everything between them
        runs on the B200 kernels, with two fusions the method-by-method drop-in cannot make: token-0 pooling of the minor
        frames happens inside the compaction (no 249 MB clone, :551 + :562), and ``embed_tokens`` + the <image>-row
        replacement (:567-572) is one pass."""
        vision_encoder = self.get_vision_encoder()
        if vision_encoder is None or pixel_values is None or input_ids.shape[1] == 1:  # :524-526 text-only
            return input_ids, attention_mask, position_ids, past_key_values, None, labels
        B, N = input_ids.shape  # :529-541
        input_ids = input_ids.reshape(B * N)
        attention_mask = attention_mask.reshape(B * N) if attention_mask is not None else None
        position_ids = position_ids.reshape(B * N) if position_ids is not None else None
        labels = labels.reshape(B * N) if labels is not None else None
        img_id = self.config.image_token_index
        mm_features = None
        if if_visual:
            batched_num_patches = grid_sizes.prod(dim=1).div(merge_sizes ** 2).long()  # :545
            mm_features = self.encode_images(pixel_values, grid_sizes, merge_sizes)  # :546 (untouched encoder + projector)
            mm_features = self._get_valid_visual_tokens(mm_features, batched_num_patches, modals)  # :547
            assert mm_features.shape[0] % total_image_num == 0, \
                f"mm_features.shape[0] % patch_num != 0, {mm_features.shape[0]} % {total_image_num} != 0"  # :549
            frame_indices = self.select_events_based_on_summary(mm_features, total_image_num, self.all_timestamps)  # :550
            n_image_ids = None
            if position_ids is not None:
                n_image_ids = int(input_ids.eq(img_id).sum())
            no_truncation = position_ids is None or mm_features.shape[0] == n_image_ids  # :358-359
            if self.config.use_token_compression and no_truncation:
                assert B == 1, "Token compression is only supported for batch_size=1"  # :561
                feats = _frames_bf16(mm_features)
                T = int(total_image_num)
                P = feats.shape[0] // T
                mask = self._get_compression_mask(pixel_values, batched_num_patches, grid_sizes, merge_sizes, modals,
                                                  minor_frame_indices=frame_indices)  # :552
                mask_u8 = mask.contiguous().view(torch.uint8)
                minor_flags = torch.zeros(T, dtype=torch.uint8, device=feats.device)
                if len(frame_indices):
                    minor_flags[torch.tensor(list(frame_indices), dtype=torch.int64, device=feats.device)] = 1
                # :551 + :562: x[mask] where token 0 of a minor frame is emitted as the frame's pooled mean, + :459-474
                if position_ids is None:
                    b = ops.compact_stream(feats, T, P, mask_u8, minor_flags, finalize=False, input_ids=input_ids.contiguous(),
                                           image_token_index=img_id, attention_mask=attention_mask, labels=labels)
                    out, n_kept, n_out = b["out"], b["n_kept"], b["n_out"]
                    o_ids, o_am, o_lb, o_ps = b["out_ids"], b.get("out_attention"), b.get("out_labels"), None
                else:
                    out, _, n_kept = ops.compact_tokens(feats, T, P, mask_u8, minor_flags)
                    o_ids, o_am, o_lb, o_ps, n_out = ops.filter_ids(input_ids.contiguous(), img_id, mask_u8, attention_mask, labels,
                                                                    position_ids)
                nk = int(n_kept.item())
                no, n_img = (int(v) for v in n_out.tolist())
                if n_img != mask_u8.numel():  # :463 raises in the reference
                    raise RuntimeError(f"shape mismatch: value tensor of shape [{mask_u8.numel()}] cannot be broadcast to indexing "
                                       f"result of shape [{n_img}]")
                mm_features = out[:nk] if out.dtype == mm_features.dtype else out[:nk].to(mm_features.dtype)
                input_ids = o_ids[:no]
                attention_mask = None if o_am is None else o_am[:no].to(attention_mask.dtype)
                labels = None if o_lb is None else o_lb[:no].to(labels.dtype)
                position_ids = None if o_ps is None else o_ps[:no].to(position_ids.dtype)
            else:  # truncated visual tokens / compression switched off: the same kernels, method by method (:551-564)
                mm_features = self.compress_unimportant_events(mm_features, mm_features.shape[0] // total_image_num,
                                                               frame_indices)
                compression_mask = self._get_compression_mask(pixel_values, batched_num_patches, grid_sizes, merge_sizes,
                                                              modals, minor_frame_indices=frame_indices)
                mm_features, compression_mask = self._maybe_truncate_visual_tokens(
                    mm_features, compression_mask, batched_num_patches, modals, input_ids, position_ids)
                if self.config.use_token_compression:
                    assert B == 1, "Token compression is only supported for batch_size=1"
                    mm_features, input_ids, attention_mask, position_ids, labels = self._compress_visual_tokens(
                        compression_mask, mm_features, input_ids, attention_mask, position_ids, labels)
        # 4./5. embed text tokens, replace the <image> rows (:567-572): one pass
        emb_w = _embed_weight(self.get_model(), None if mm_features is None else mm_features.shape[1])
        if not input_ids.is_cuda:
            raise RuntimeError("cogstream_b200: input_ids are on the CPU; there is no CPU fallback")
        inputs_embeds, _ = ops.embed_scatter(input_ids.contiguous(), img_id if if_visual else -1, emb_w,
                                             mm_features.contiguous() if if_visual else None)
        C = inputs_embeds.shape[-1]  # :575-583
        inputs_embeds = inputs_embeds.reshape(B, -1, C)
        attention_mask = attention_mask.view(B, -1) if attention_mask is not None else None
        labels = labels.view(B, -1) if labels is not None else None
        position_ids = position_ids.view(B, -1) if position_ids is not None else None
        return None, attention_mask, position_ids, past_key_values, inputs_embeds, labels


_METHODS = ("select_events_based_on_summary", "compress_unimportant_events", "_get_compression_mask",
            "_compress_visual_tokens")


def install(chat_module, *, model_class_name: str = "Videollama3MetaForCausalLM", call_site: bool = True):
    """Rebind the hot path of an imported reference ``model.cogreasoner_chat`` module onto the B200 kernels.
    ``call_site=True`` also replaces ``prepare_inputs_labels_for_multimodal`` (:509-585) by the fused version; with
    ``False`` only the four methods and two functions it calls are swapped."""
    chat_module.kmeans_with_time_min_max = kmeans_with_time_min_max  # imported at cogreasoner_chat.py:32
    chat_module.select_additional_frames = select_additional_frames
    chat_module.create_visual_summary_prompt = create_visual_summary_prompt
    cls = getattr(chat_module, model_class_name)
    for name in _METHODS + (("prepare_inputs_labels_for_multimodal",) if call_site else ()):
        setattr(cls, name, getattr(VisualStreamCompressionMixin, name))
    for name in ("csb_frames_per_event", "csb_min_events", "csb_n_representatives", "csb_relevance_threshold"):
        setattr(cls, name, getattr(VisualStreamCompressionMixin, name))
    return cls
