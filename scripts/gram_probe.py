import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import ops, synthetic
T, P, D = int(os.environ.get('T', 600)), int(os.environ.get('P', 180)), 1152
x = synthetic.make_features(T, P, D, 0)[0].view(T, -1).cuda()
G = torch.empty(T, T, dtype=torch.float64, device="cuda")
ws = torch.empty(ops.gram_workspace_bytes(T, P * D), dtype=torch.uint8, device="cuda")
# NOTE: these are wall-clock times of back-to-back calls.  A full call is GPU-bound (Gram kernel ~80 us + reduce ~6 us on the stream);
# the "reduce kernel only" line is HOST-bound (~15 us per call: tensor-map encode + two launches), it says nothing about the
# reduce kernel, which takes ~6 us of GPU time (measured by replacing it with an empty kernel: full 86.2 -> 80.4 us).
for name, impl in [("full", 0), ("no-mma (loads only)", 16), ("no-tma (mma only)", 32), ("reduce kernel only (host-bound)", 256), ("full", 0)]:
    for _ in range(3): ops.gram(x, impl=impl, out=G, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.gram(x, impl=impl, out=G, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:32s} {e0.elapsed_time(e1)/20*1e3:7.1f} us per call (Gram kernel + ~6 us reduce kernel)")

# per-CTA time marks (impl bit 7): where the kernel time goes
import numpy as np
n_units, n_kr = int(os.environ.get('NU', 4)), int(os.environ.get('NKR', 37))
wsz = ops.gram_workspace_bytes(T, P * D)
ws2 = torch.zeros(wsz + n_units * n_kr * 64, dtype=torch.uint8, device="cuda")
for _ in range(3): ops.gram(x, impl=128, out=G, workspace=ws2)
torch.cuda.synchronize()
m = ws2[wsz:].view(torch.int64).view(n_kr * n_units, 8).cpu().numpy().astype(np.int64)
t0 = m[:, 0].min()
names = ["entry", "setup done", "first operands", "last MMA issued", "accum complete", "drain done", "last TMA issued"]
print("marks relative to the first CTA's entry, us: min / median / max over the 148 CTAs")
for i, n in enumerate(names):
    v = (m[:, i] - t0) / 1e3
    print(f"  {n:18s} {v.min():7.2f} {np.median(v):7.2f} {v.max():7.2f}")
for u in range(min(n_units, 4)):
    rows = m[u::n_units] if True else None
    v = (rows[:, 4] - rows[:, 2]) / 1e3
    print(f"  unit {u}: mainloop (first operands -> accum complete) median {np.median(v):.2f} us, max {v.max():.2f}")
life = (m[:, 5] - m[:, 0]) / 1e3
main = (m[:, 3] - m[:, 2]) / 1e3
print(f"  CTA lifetime median {np.median(life):.1f} us (min {life.min():.1f}, max {life.max():.1f}); first operands -> last MMA issued median {np.median(main):.1f} us; drain after last accum median {np.median((m[:,5]-m[:,4])/1e3):.1f} us")
print("  distinct SMs:", len(set(m[:, 7].tolist())))
