"""Full-size (BASELINE config 2) k-means parity: GPU (tcgen05 / CUDA-core Gram) vs the CPU oracle, several seeds."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import kmeans_with_time_min_max, synthetic
from oracle import cogstream_oracle as orc
torch.set_num_threads(os.cpu_count())
T, P, D = 600, int(os.environ.get("P", 180)), 1152
k = math.ceil(T / 15)
dev = torch.device("cuda:0")
for seed in range(int(os.environ.get("SEEDS", 6))):
    feats, _ = synthetic.make_features(T, P, D, seed)
    ts = torch.arange(T, dtype=torch.float64)
    g = torch.Generator().manual_seed(seed)
    init = torch.randperm(T, generator=g)[:k].tolist()
    pool = torch.randint(0, T, (64,), generator=g).tolist()
    t0 = time.time()
    info = orc.KMeansInfo()
    c, ct, a = orc.kmeans_with_time_min_max(feats.view(T, P, D), ts, k, init_indices=init, reseed_indices=pool, info=info)
    t_cpu = time.time() - t0
    out = []
    for impl in (0, 1):
        cg, ctg, ag, ig = kmeans_with_time_min_max(feats.view(T, P, D).to(dev), ts.to(dev), k, init_indices=init, reseed_indices=pool,
                                                   gram_impl=impl, return_info=True)
        nd = int((ag.cpu() != a).sum())
        out.append(f"impl{impl}: ndiff={nd} iters={ig.iterations} margin={ig.min_margin:.2e} cent_rel={float((cg.cpu() - c).norm() / c.norm()):.1e}")
    print(f"seed {seed}: cpu iters={info.iterations} margin={info.min_margin:.2e} ({t_cpu:.1f}s) | " + " | ".join(out), flush=True)
