import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import ops, synthetic
T, P = 600, 180
pix = synthetic.make_pixels(T, P, 2, 588, 0).cuda()
W = pix.numel() // (T * P)
ref = None
for cfg in ["old", "8x4", "12x3", "16x2", "16x3", "4x8"]:
    if cfg == "old":
        os.environ["CSB_PIX_CFG"] = "1x999"
    else:
        os.environ["CSB_PIX_CFG"] = cfg
    out = torch.empty(T * P, dtype=torch.uint8, device="cuda")
    for _ in range(3): ops.pixel_diff(pix, T, P, W, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.pixel_diff(pix, T, P, W, out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    if ref is None: ref = out.clone()
    print(f"{cfg:6s} {us:7.1f} us  {pix.numel()*2/us/1e3:7.1f} GB/s  same={torch.equal(out, ref)}")
