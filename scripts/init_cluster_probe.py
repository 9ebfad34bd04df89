"""k-means++ of long streams: the thread-block-cluster form (hardware barrier + DSMEM exchange per sampling round) next to the
cooperative-grid form (software barrier through the L2), same Gram matrix, same generator state / same injected noise.
Prints one JSON line per stream length: init-only and full-solve times of both forms and whether they picked the same centres.
Usage (one B200): python scripts/init_cluster_probe.py [T ...]      (default: 2300 4100 8000 14400)"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import ops, synthetic

dev = torch.device("cuda:0")
P, D = 16, 1152
sizes = [int(v) for v in sys.argv[1:]] or [2300, 4100, 8000, 14400]


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for T in sizes:
    k = math.ceil(T / 15)
    feats, _ = synthetic.make_features(T, P, D, 0)
    x2 = feats.view(T, -1).to(dev)
    ts = (torch.arange(T, dtype=torch.float64) / 2.0).to(dev)
    G = ops.gram(x2)
    buf = ops.KMeansBuffers(T, k, torch.float64, dev)
    rng = torch.tensor([1234, 0], dtype=torch.int64, device=dev)
    g = torch.Generator().manual_seed(T)
    noise = torch.empty(k - 1, T).exponential_(1, generator=g).to(dev)
    pool = torch.randint(0, T, (4096,), generator=g).to(dev)
    first = int(torch.randint(0, T, (1,), generator=g))
    row = {"T": T, "k": k}
    picks = {}
    for mode in ("cluster", "grid"):
        os.environ["CSB_INIT_CLUSTER"] = "1" if mode == "cluster" else "0"

        def solve_rng(max_it):
            rng.copy_(torch.tensor([1234, 0], dtype=torch.int64))
            ops.kmeans_solve(G, ts, buf, n_init_given=0, exp_noise=None, reseed_pool=None, max_iteration=max_it, rng_state=rng)

        def solve_noise(max_it):
            buf.init_indices[0] = first
            ops.kmeans_solve(G, ts, buf, n_init_given=1, exp_noise=noise, reseed_pool=pool, max_iteration=max_it)

        row[f"{mode}_init_rng_ms"] = round(timed(lambda: solve_rng(0), 3), 4)
        picks[mode, "rng"] = buf.init_indices.cpu().clone()
        row[f"{mode}_full_rng_ms"] = round(timed(lambda: solve_rng(30), 3), 4)
        picks[mode, "rng_assign"] = buf.assignments.cpu().clone()
        st = ops.decode_status(buf.status.cpu())
        row[f"{mode}_iterations"] = st["iterations"]
        row[f"{mode}_init_noise_ms"] = round(timed(lambda: solve_noise(0), 3), 4)
        picks[mode, "noise"] = buf.init_indices.cpu().clone()
    # per-phase clock cycles of a sampling round (thread 0 of CTA 0): loads + update | block reduction + publish | barrier |
    # peers + scores + broadcast | rest
    os.environ["CSB_INIT_PROFILE"] = "1"
    off = (8 * T + 255) // 256 * 256
    for mode, csz, blk in (("cluster", None, None), ("cluster_nopush", None, None), ("cluster", None, 512), ("cluster", 8, None),
                           ("cluster", 8, 512), ("grid", None, None)):
        os.environ["CSB_INIT_CLUSTER"] = "1" if mode.startswith("cluster") else "0"
        os.environ["CSB_INIT_PUSH"] = "0" if mode.endswith("nopush") else "1"
        if csz:
            os.environ["CSB_INIT_CLUSTER_SIZE"] = str(csz)
        if blk:
            os.environ["CSB_INIT_CLUSTER_BLOCK"] = str(blk)
        buf.init_indices[0] = first
        ops.kmeans_solve(G, ts, buf, n_init_given=1, exp_noise=noise, reseed_pool=pool, max_iteration=-1)
        torch.cuda.synchronize()
        tag = f"{mode}{csz or ''}{'_b512' if blk else ''}"
        row[f"{tag}_phase_cycles"] = [round(float(v), 1) for v in buf.workspace[off: off + 24].view(torch.float32).cpu()]
        picks[tag] = buf.init_indices.cpu().clone()
        ms = timed(lambda: ops.kmeans_solve(G, ts, buf, n_init_given=1, exp_noise=noise, reseed_pool=pool, max_iteration=-1), 2)
        row[f"{tag}_prof_ms"] = round(ms, 4)
        os.environ.pop("CSB_INIT_CLUSTER_SIZE", None)
        os.environ.pop("CSB_INIT_CLUSTER_BLOCK", None)
    row["same_centres_variants"] = all(bool(torch.equal(picks["grid"], picks[t])) for t in ("cluster", "cluster_nopush", "cluster_b512", "cluster8", "cluster8_b512"))
    os.environ.pop("CSB_INIT_PROFILE")
    os.environ.pop("CSB_INIT_PUSH")
    os.environ.pop("CSB_INIT_CLUSTER")
    row["same_centres_rng"] = bool(torch.equal(picks["cluster", "rng"], picks["grid", "rng"]))
    row["same_assignments_rng"] = bool(torch.equal(picks["cluster", "rng_assign"], picks["grid", "rng_assign"]))
    row["same_centres_noise"] = bool(torch.equal(picks["cluster", "noise"], picks["grid", "noise"]))
    row["distinct_centres"] = int(torch.unique(picks["cluster", "rng"]).numel())
    print(json.dumps(row), flush=True)
    del G, x2, buf, noise
    torch.cuda.empty_cache()
