"""Timeline of the first Lloyd iterations at BASELINE config 4 under the assignment-pass switches (CSB_ASSIGN_EXACT: 0 = filter,
1 = every cluster exact, 2 / 3 = timing aids without the exact pass / without the estimate pass)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import ops, synthetic

T, P, D = int(os.environ.get("T", 14400)), 16, 1152
k = math.ceil(T / 15)
dev = torch.device("cuda:0")
feats, _ = synthetic.make_features(T, P, D, 0)
x2 = feats.view(T, -1).to(dev)
ts = (torch.arange(T, dtype=torch.float64) / 2.0).to(dev)
G = ops.gram(x2)
buf = ops.KMeansBuffers(T, k, torch.float64, dev)
g = torch.Generator().manual_seed(0)
init = torch.randperm(T, generator=g)[:k].to(dev)
pool = torch.randint(0, T, (4096,), generator=g).to(dev)
names = {1: "start", 2: "init done", 3: "S0 done", 10: "assign", 11: "sync", 12: "rebuild", 13: "S", 14: "sync", 15: "shift"}
for mode in sys.argv[1:] or ["0", "1", "2", "3"]:
    os.environ["CSB_ASSIGN_EXACT"] = mode
    if mode == "0":
        os.environ.pop("CSB_ASSIGN_EXACT")
    for rep in range(2):
        buf.init_indices.copy_(init)
        ops.kmeans_solve(G, ts, buf, n_init_given=k, exp_noise=None, reseed_pool=pool, max_iteration=2 if mode in "01" else 1)
        torch.cuda.synchronize()
    prev, out = 0, []
    for tag, t in buf.timeline():
        out.append(f"{names.get(tag, tag)} {(t - prev) / 1e3:.1f}")
        prev = t
    print(f"mode {mode}: " + " | ".join(out), flush=True)
