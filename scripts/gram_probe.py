import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import ops, synthetic
T, P, D = 600, 180, 1152
x = synthetic.make_features(T, P, D, 0)[0].view(T, -1).cuda()
G = torch.empty(T, T, dtype=torch.float64, device="cuda")
ws = torch.empty(ops.gram_workspace_bytes(T, P * D), dtype=torch.uint8, device="cuda")
for name, impl in [("full", 0), ("no-mma (loads only)", 16), ("no-tma (mma only)", 32), ("mma only, N=128 only", 32 + 64), ("full, N=128 only", 64), ("full", 0)]:
    for _ in range(3): ops.gram(x, impl=impl, out=G, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.gram(x, impl=impl, out=G, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:22s} {e0.elapsed_time(e1)/20*1e3:7.1f} us (incl. ~13 us reduce kernel)")
