"""k-means++ of mid-size streams (1024 < T <= 2048 takes the single-CTA kernel, longer streams the thread-block cluster): both
forms at the same sizes (CSB_INIT_GRID=0 / 1 force either), built-in generator, init-only solves."""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import ops, synthetic

dev = torch.device("cuda:0")
for T in [int(v) for v in sys.argv[1:]] or [700, 1100, 1500, 2000, 2300, 3000]:
    k = math.ceil(T / 15)
    feats, _ = synthetic.make_features(T, 4, 256, 0)
    G = ops.gram(feats.view(T, -1).to(dev))
    ts = torch.arange(T, dtype=torch.float64, device=dev)
    buf = ops.KMeansBuffers(T, k, torch.float64, dev)
    rng0 = torch.tensor([99, 0], dtype=torch.int64, device=dev)
    rng = rng0.clone()
    row = {"T": T, "k": k}
    picks = {}
    for mode in ("single_cta", "cluster"):
        os.environ["CSB_INIT_GRID"] = "0" if mode == "single_cta" else "1"
        if T <= 1024:
            os.environ["CSB_FUSED_INIT"] = "0"

        def solve(max_it):
            rng.copy_(rng0)
            ops.kmeans_solve(G, ts, buf, n_init_given=0, exp_noise=None, reseed_pool=None, max_iteration=max_it, rng_state=rng)

        solve(-1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            solve(-1)
        e1.record(); torch.cuda.synchronize()
        row[f"{mode}_init_ms"] = round(e0.elapsed_time(e1) / 5, 4)
        row[f"{mode}_us_per_round"] = round(1e3 * e0.elapsed_time(e1) / 5 / (k - 1), 3)
        picks[mode] = buf.init_indices.cpu().clone()
    os.environ.pop("CSB_INIT_GRID"); os.environ.pop("CSB_FUSED_INIT", None)
    row["same_centres"] = bool(torch.equal(picks["single_cta"], picks["cluster"]))
    print(json.dumps(row), flush=True)
