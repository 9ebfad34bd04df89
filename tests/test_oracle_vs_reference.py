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
