"""CPU: the C-ABI library exports exactly what include/cogstream_b200.h declares; host-side logic."""
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT
from cogstream_b200 import STATS_FIELDS, StreamConfig, _lib, ops, pad_rows, shard_streams, synthetic
from cogstream_b200.compression import VisualStreamCompressionMixin, create_visual_summary_prompt
from cogstream_b200.kmeans_with_time import kmeans_with_time_min_max


def _declared():
    text = open(os.path.join(ROOT, "include", "cogstream_b200.h")).read()
    return sorted(set(re.findall(r"CSB_API\s+[\w\s\*]+?\b(csb_\w+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    declared = _declared()
    assert len(declared) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\sT\s+(csb_\w+)", out)))
    assert exported == declared
    assert sorted(_lib.PROTOTYPES) == declared  # the ctypes table mirrors the header one to one
    for name in declared:
        assert hasattr(built_lib, name)


def test_abi_version_and_sizes(built_lib):
    assert built_lib.csb_abi_version() == 2
    # pure host arithmetic (no GPU): workspace sizes are positive and monotone
    a = built_lib.csb_gram_workspace_bytes(600, 207360)
    b = built_lib.csb_gram_workspace_bytes(64, 4608)
    assert a > 0 and b > 0
    assert built_lib.csb_kmeans_workspace_bytes(600, 40) > 2 * 600 * 40 * 8
    assert built_lib.csb_compact_workspace_bytes(108000, 108037) >= 4 * 108037
    assert built_lib.csb_gram_workspace_bytes(0, 0) == 0


def test_sass_is_blackwell_native(built_lib):
    """The Gram kernel must be tcgen05 + TMA + TMEM (UTCHMMA / UTMALDG / LDTM in SASS), not mma.sync."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")
    assert "sm_100a" in sass or "EF_CUDA_SM100" in sass


def test_no_cpu_fallback():
    x = torch.zeros(4, 2, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kmeans_with_time_min_max(x, torch.arange(4.0), 2)
    with pytest.raises(RuntimeError, match="CPU"):
        ops.gram(x.view(4, 16))
    with pytest.raises(RuntimeError, match="CPU"):
        ops.pixel_mask(torch.zeros(8, 4, dtype=torch.bfloat16), 2, 2, 8)

    class S(VisualStreamCompressionMixin):
        pass

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S().compress_unimportant_events(torch.zeros(8, 8, dtype=torch.bfloat16), 2, [1])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cogstream_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_gate_and_cluster_count():
    cfg = StreamConfig()
    assert cfg.clusters_for(600) == 40 and cfg.gate_open(600)
    assert cfg.clusters_for(135) == 9 and not cfg.gate_open(135)  # k <= 9 -> no compression (cogreasoner_chat.py:281)
    assert cfg.clusters_for(136) == 10 and cfg.gate_open(136)
    assert StreamConfig(cluster_num=8).gate_open(64) and not StreamConfig(cluster_num=8).gate_open(8)


def test_shard_and_pad():
    assert shard_streams(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((shard_streams(256, r, 8) for r in range(8)), [])) == list(range(256))
    F = len(STATS_FIELDS)
    rows = pad_rows([torch.arange(F), torch.arange(F) + 1], 3, "cpu")
    assert rows.shape == (3, F) and rows[2].sum() == 0 and rows[1, 0] == 1
    assert "reseed_overflow" in STATS_FIELDS


def test_prompt_layout():
    p = create_visual_summary_prompt(6, [0.0, 12.34, 99.0])
    assert p.count("<image>") == 6 and "Time 12.3s:<image><image>," in p and p.endswith("<|im_start|>assistant")


def test_synthetic_is_deterministic():
    a = synthetic.make_stream(40, 3, 16, seed=5, C=4)
    b = synthetic.make_stream(40, 3, 16, seed=5, C=4)
    assert torch.equal(a.features, b.features) and torch.equal(a.pixel_values, b.pixel_values)
    assert a.features.dtype == torch.bfloat16 and a.timestamps.dtype == torch.float64
    assert a.pixel_values.shape == (40 * 3 * 4, 4) and a.input_ids.eq(synthetic.IMAGE_TOKEN_INDEX).sum() == 120


def test_custom_ops_registered_and_shape_inference():
    """The torch custom-op layer exists, infers shapes on meta tensors and has no CPU kernel."""
    import cogstream_b200  # noqa: F401

    names = ("gram", "gram_extend", "kmeans_solve", "centroid_update", "centroid_gather", "dist_to_assigned", "dist_to_assigned_gram",
             "select_representatives", "representatives_gram", "event_order", "gather_rows", "cosine_flags", "minor_frames",
             "relevance_minor", "relevance_select", "relevance_select_hidden", "pool_norm", "pool_frames_inplace", "pixel_diff",
             "mask_finalize", "pixel_mask", "compact_tokens", "filter_ids", "compact_stream", "embed_scatter", "embed_scatter_ordered",
             "seq_mean_pool", "ln_merge")
    for name in names:
        assert hasattr(torch.ops.cogstream_b200, name), name
    # every compute entry point of the header is reachable through the dispatcher (plumbing / workspace-size queries aside)
    covered = {"csb_" + n.replace("_hidden", "").replace("embed_scatter_ordered", "embed_scatter_ordered") for n in names}
    plumbing = {"csb_abi_version", "csb_last_error", "csb_check_device", "csb_kernel_launches", "csb_pixel_diff_ex",
                "csb_kmeans_lloyd_phase", "csb_kmeans_lloyd_state_bytes"}  # (driven by cogstream_b200.sharded, a multi-launch protocol)
    missing = [d for d in _declared() if d not in covered and d not in plumbing and not d.endswith("_workspace_bytes")]
    assert missing == [], missing
    m = torch.empty(12, dtype=torch.uint8, device="meta")
    ids = torch.empty(20, dtype=torch.int64, device="meta")
    o = torch.ops.cogstream_b200.compact_stream(torch.empty(12, 64, dtype=torch.bfloat16, device="meta"), 6, 2, m, None, ids, 7)
    assert o[0].shape == (12, 64) and o[3].shape == (20,) and o[5].shape == (2,)
    x = torch.empty(10, 64, dtype=torch.bfloat16, device="meta")
    assert torch.ops.cogstream_b200.gram(x).shape == (10, 10)
    out, kept, n = torch.ops.cogstream_b200.compact_tokens(x, 5, 2, torch.empty(10, dtype=torch.uint8, device="meta"))
    assert out.shape == (10, 64) and kept.dtype == torch.int32 and n.shape == (1,)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.cogstream_b200.gram(torch.zeros(4, 8, dtype=torch.bfloat16))


def test_install_rebinds_reference_names():
    """install() swaps exactly the names the reference's hot path resolves at call time (cogreasoner_chat.py:32, :50, :93,
    :278, :383, :434, :449 and, with call_site=True, :509)."""
    import types
    from cogstream_b200 import compression, kmeans_with_time

    def fake_module():
        mod = types.ModuleType("fake_chat")

        class Videollama3MetaForCausalLM:
            def select_events_based_on_summary(self): return "ref"
            def compress_unimportant_events(self): return "ref"
            def _get_compression_mask(self): return "ref"
            def _compress_visual_tokens(self): return "ref"
            def prepare_inputs_labels_for_multimodal(self): return "ref"
            def _maybe_truncate_visual_tokens(self): return "ref"

        mod.Videollama3MetaForCausalLM = Videollama3MetaForCausalLM
        mod.kmeans_with_time_min_max = mod.select_additional_frames = mod.create_visual_summary_prompt = lambda *a: "ref"
        return mod

    mod = fake_module()
    cls = compression.install(mod)
    assert mod.kmeans_with_time_min_max is kmeans_with_time.kmeans_with_time_min_max
    assert mod.select_additional_frames is compression.select_additional_frames
    for name in ("select_events_based_on_summary", "compress_unimportant_events", "_get_compression_mask",
                 "_compress_visual_tokens", "prepare_inputs_labels_for_multimodal"):
        assert getattr(cls, name) is getattr(compression.VisualStreamCompressionMixin, name)
    assert cls._maybe_truncate_visual_tokens(None) == "ref"  # not on the path: untouched
    mod2 = fake_module()
    cls2 = compression.install(mod2, call_site=False)
    assert cls2.prepare_inputs_labels_for_multimodal(None) == "ref"
    import inspect
    sig = inspect.signature(compression.VisualStreamCompressionMixin.prepare_inputs_labels_for_multimodal)
    assert list(sig.parameters) == ["self", "input_ids", "attention_mask", "position_ids", "past_key_values", "labels",
                                    "pixel_values", "grid_sizes", "merge_sizes", "modals", "total_image_num", "if_visual"]


def test_frame_range_partition():
    from cogstream_b200 import frame_range
    for T, world in ((14400, 8), (600, 7), (130, 8), (5, 8), (1000, 1)):
        seen = []
        for r in range(world):
            per, b, e = frame_range(T, r, world)
            assert per * world >= T and 0 <= b <= e <= T and e - b <= per and (b % per == 0 or b == e)
            seen += list(range(b, e))
        assert seen == list(range(T))
