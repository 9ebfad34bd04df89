"""Accuracy of the tcgen05 Gram matrix at full config-2 size: entries and the squared distances derived from them."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import ops, synthetic
T, P, D = 600, 180, 1152
for iid in (False, True):
    x = synthetic.make_features(T, P, D, 0, iid=iid)[0].view(T, -1).cuda()
    xd = x.double()
    ref = xd @ xd.T
    dg = ref.diagonal()
    d2_ref = dg[:, None] + dg[None, :] - 2 * ref
    off = ~torch.eye(T, dtype=torch.bool, device="cuda")
    for name, impl in (("tcgen05", 0), ("cuda-core", 1)):
        G = ops.gram(x, impl=impl)
        e = (G - ref).abs()
        d2 = G.diagonal()[:, None] + G.diagonal()[None, :] - 2 * G
        rel = ((d2 - d2_ref).abs() / d2_ref.clamp_min(1e-9))[off]
        sgn = ((G - ref) / ref.abs().clamp_min(1.0))
        print(f"iid={iid} {name:9s}: max|dG|/max diag {float(e.max() / dg.max()):.2e}  mean signed rel err diag {float(sgn.diagonal().mean()):+.2e} "
              f" d2 rel err: max {float(rel.max()):.2e} median {float(rel.median()):.2e}  (min d2 {float(d2_ref[off].min()):.1f})")
