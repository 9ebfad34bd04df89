"""BASELINE config 5: k in {8, 32, 128} x token budget {1k, 4k, 16k} (T = 600 -> P in {2, 7, 27}), one stream on one GPU, next
to the CPU oracle (the reference's PyTorch CPU path restated) on the box's host cores.  Writes one JSON object per line."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import StreamCompressor, StreamConfig, synthetic
from oracle import cogstream_oracle as orc  # CPU baseline arm only (never on the product path)

dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count())
T, D = 600, 1152
rows = []
for k in (8, 32, 128):
    for P, budget in ((2, "1k"), (7, "4k"), (27, "16k")):
        st = synthetic.make_stream(T, P, D, seed=k + P, C=48)
        ev, q = synthetic.make_embeddings(k, 3584, seed=k + P)
        d = st.to(dev); evd, qd = ev.to(dev), q.to(dev)
        W = st.pixel_values.numel() // (T * P)
        sc = StreamCompressor(T, P, D, W, st.input_ids.numel(), device=dev, cfg=StreamConfig(cluster_num=k))
        g = torch.Generator().manual_seed(k * 100 + P)
        noise = torch.empty(k - 1, T).exponential_(1, generator=g)
        first = int(torch.randint(0, T, (1,), generator=g))
        pool = [int(v) for v in torch.randint(0, T, (4 * k,), generator=g)]
        kw = dict(first_index=first, exp_noise=noise.to(dev), reseed_pool=torch.tensor(pool, dtype=torch.int64, device=dev))
        def step():
            sc.phase1(d.features, d.timestamps, **kw)
            sc.phase2(d.features, evd, qd, d.pixel_values, d.input_ids, d.attention_mask)
        for _ in range(3): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): step()
        e1.record(); torch.cuda.synchronize()
        gpu_ms = e0.elapsed_time(e1) / 10
        r = sc.results()
        # CPU: same inputs, same injected randomness
        import random
        orig = random.randint
        t0 = time.perf_counter()
        try:
            random.randint = lambda a, b: first
            p1 = orc.phase1(st.features, T, st.timestamps, cluster_num=k, exp_noise=noise, reseed_indices=pool, min_clusters=0)
        finally:
            random.randint = orig
        p2 = orc.phase2(st.features, p1, ev, q, st.pixel_values, st.grid_sizes, st.merge_sizes, st.modals, st.input_ids,
                        synthetic.IMAGE_TOKEN_INDEX, st.attention_mask)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        same = torch.equal(r["assignments"].cpu(), p1.assignments) and torch.equal(r["mask"].cpu(), p2.mask) and r["minor"] == p2.minor
        row = dict(T=T, P=P, budget=budget, k=k, gpu_ms_per_stream=round(gpu_ms, 4), gpu_frames_per_s=round(T / gpu_ms * 1e3),
                   cpu_ms_per_stream=round(cpu_ms, 1), cpu_frames_per_s=round(T / cpu_ms * 1e3), cpu_cores=os.cpu_count(),
                   iterations=r["status"]["iterations"], n_kept=int(r["n_kept"]), identical_to_cpu=bool(same))
        rows.append(row)
        print(json.dumps(row), flush=True)
