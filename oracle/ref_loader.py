"""Load the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY (see oracle/cogstream_oracle.py header).  ``/root/reference`` does
not exist on the GPU box; callers must check :func:`available` and skip.

Recipe verified in SURVEY.md Appendix A:
  * ``model/kmeans_with_time.py`` needs only torch + random -> loaded by file path;
  * ``model/cogreasoner_chat.py`` imports as the namespace package ``model`` once
    ``/root/reference`` is on ``sys.path`` (transformers 5.5 present).
Init injection ("same inputs and same init"): the names ``random`` and ``torch`` inside the
loaded k-means module are rebound to small proxies for the duration of one call, so the
reference's own lines :41, :57, :60, :118 replay a prescribed sequence.  No reference
source is modified or copied.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys

import torch
import types

REFERENCE_ROOT = os.environ.get("COGSTREAM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "kmeans_with_time.py"))


_cache = {}


def kmeans_module():
    if "km" not in _cache:
        path = os.path.join(REFERENCE_ROOT, "model", "kmeans_with_time.py")
        spec = importlib.util.spec_from_file_location("cogstream_ref_kmeans", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _cache["km"] = mod
    return _cache["km"]


def chat_module():
    if "cc" not in _cache:
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import model.cogreasoner_chat as cc  # noqa: WPS433 (namespace package under the reference root)

        _cache["cc"] = cc
    return _cache["cc"]


def encoder_module():
    """The reference's ``model.modeling_videollama3_encoder`` (imports in the build container: flash_attn is present)."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    return importlib.import_module("model.modeling_videollama3_encoder")


def ref_encoder_tail(hidden_states, grid_sizes, merge_sizes, post_layernorm):
    """``Videollama3VisionEncoderModel.forward`` (:477-506) run UNMODIFIED on a stand-in whose embeddings / transformer stack are the
    identity, i.e. exactly its tail: post_layernorm + the per-item bilinear merge."""
    import types
    M = encoder_module().Videollama3VisionEncoderModel
    stub = types.SimpleNamespace(embeddings=lambda x: x, encoder=lambda h, g, m: h, post_layernorm=post_layernorm)
    with torch.no_grad():
        return M.forward(stub, hidden_states, grid_sizes, merge_sizes)


def stub_self(image_token_index: int = 151665):
    return types.SimpleNamespace(
        config=types.SimpleNamespace(image_token_index=image_token_index, use_token_compression=True)
    )


class _ReplayRandom:
    def __init__(self, values):
        self._it = iter(values)
        self.consumed = []

    def randint(self, a, b):
        v = int(next(self._it))
        assert a <= v <= b
        self.consumed.append(v)
        return v


class _TorchProxy:
    """Forwards everything to torch except ``multinomial``, which replays injected centres or
    evaluates argmax(probs / exp_noise[round]) (what ATen's multinomial computes for one draw)."""

    def __init__(self, torch_mod, picks=None, exp_noise=None):
        self._t = torch_mod
        self._picks = iter(picks) if picks is not None else None
        self._noise = exp_noise
        self._round = 0

    def __getattr__(self, name):
        return getattr(self._t, name)

    def multinomial(self, probs, n):
        assert n == 1
        if self._picks is not None:
            v = int(next(self._picks))
            return self._t.tensor([v])
        r = self._round
        self._round += 1
        return self._t.argmax(probs / self._noise[r]).reshape(1)


@contextlib.contextmanager
def injected_rng(randint_values, picks=None, exp_noise=None):
    """Rebind ``random`` / ``torch`` inside the reference k-means module for one call."""
    import torch

    km = kmeans_module()
    old_r, old_t = km.random, km.torch
    rr = _ReplayRandom(randint_values)
    km.random = rr
    if picks is not None or exp_noise is not None:
        km.torch = _TorchProxy(torch, picks, exp_noise)
    try:
        yield rr
    finally:
        km.random, km.torch = old_r, old_t


def ref_kmeans(features, timestamp, k, *, init_indices=None, exp_noise=None, first=None, reseed_indices=(), **kw):
    """Run the reference k-means with an injected init.

    init_indices: all k centres (first via randint, the rest via multinomial replay).
    exp_noise + first: first centre via randint replay, the rest sampled with the given noise.
    reseed_indices: values for the empty-cluster randint calls.
    """
    km = kmeans_module()
    if init_indices is not None:
        rvals = [init_indices[0], *reseed_indices]
        with injected_rng(rvals, picks=list(init_indices[1:])) as rr:
            out = km.kmeans_with_time_min_max(features, timestamp, k, **kw)
    elif exp_noise is not None:
        with injected_rng([first, *reseed_indices], exp_noise=exp_noise) as rr:
            out = km.kmeans_with_time_min_max(features, timestamp, k, **kw)
    else:
        return km.kmeans_with_time_min_max(features, timestamp, k, **kw), []
    return out, rr.consumed[1:]
