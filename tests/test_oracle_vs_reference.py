"""CPU, build container only: the oracle against the live, unmodified reference on fresh seeds."""
import random

import pytest
import torch

from cogstream_b200 import synthetic
from oracle import cogstream_oracle as orc
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted (GPU box)")


@pytest.mark.parametrize("T,P,D,k,seed", [(80, 3, 48, 7, 101), (120, 2, 96, 10, 102)])
def test_kmeans_live(T, P, D, k, seed):
    st = synthetic.make_stream(T, P, D, seed, C=4)
    f3 = st.features.view(T, P, D)
    km = ref_loader.kmeans_module()
    random.seed(seed); torch.manual_seed(seed)
    r = km.kmeans_with_time_min_max(f3, st.timestamps, k)
    random.seed(seed); torch.manual_seed(seed)
    o = orc.kmeans_with_time_min_max(f3, st.timestamps, k)
    assert all(torch.equal(a, b) for a, b in zip(r, o))


def test_prompt_matches_reference():
    from cogstream_b200.compression import create_visual_summary_prompt

    cc = ref_loader.chat_module()
    ts = torch.tensor([0.0, 1.0, 7.25, 100.5], dtype=torch.float64)
    assert create_visual_summary_prompt(12, ts) == cc.create_visual_summary_prompt(12, ts)
    assert create_visual_summary_prompt(4, [3.0, 4.0]) == cc.create_visual_summary_prompt(4, [3.0, 4.0])


def test_encoder_tail_live():
    """oracle.encoder_tail == the unmodified Videollama3VisionEncoderModel.forward tail (modeling_videollama3_encoder.py:479-504)."""
    g = torch.Generator().manual_seed(9)
    grid = torch.tensor([[2, 6, 4], [1, 4, 4], [3, 2, 2]])
    merge = torch.tensor([2, 4, 1])
    C = 64
    hidden = torch.randn(int(grid.prod(dim=1).sum()), C, generator=g).to(torch.bfloat16)
    ln = torch.nn.LayerNorm(C, eps=1e-6).to(torch.bfloat16)
    with torch.no_grad():
        ln.weight.copy_(torch.randn(C, generator=g)); ln.bias.copy_(torch.randn(C, generator=g))
    want = ref_loader.ref_encoder_tail(hidden, grid, merge, ln)
    got = orc.encoder_tail(hidden, grid, merge, ln.weight.detach(), ln.bias.detach(), ln.eps)
    assert torch.equal(want, got)
