"""Functional check of the large / swept configurations of BASELINE.json (configs 4 and 5) on the GPU."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import StreamCompressor, StreamConfig, ops, synthetic

dev = torch.device("cuda:0")

def run(T, P, D, k=None, C=588, seed=0, check_gram=True):
    t0 = time.time()
    st = synthetic.make_stream(T, P, D, seed, C=C)
    kk = k if k is not None else math.ceil(T / 15)
    ev, q = synthetic.make_embeddings(kk, 3584, seed)
    d = st.to(dev)
    W = st.pixel_values.numel() // (T * P)
    sc = StreamCompressor(T, P, D, W, st.input_ids.numel(), device=dev, cfg=StreamConfig(cluster_num=k))
    torch.manual_seed(0)
    res = []
    for rep in range(2):
        torch.manual_seed(0)
        torch.cuda.synchronize(); t1 = time.time()
        sc.phase1(d.features, d.timestamps, first_index=5)
        sc.phase2(d.features, ev.to(dev), q.to(dev), d.pixel_values, d.input_ids, d.attention_mask)
        r = sc.results(); torch.cuda.synchronize(); dt = time.time() - t1
        res.append((r["assignments"].clone(), r["mask"].clone(), r["n_kept"], r["status"]))
    same = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    msg = f"T={T} P={P} D={D} k={kk}: {dt*1e3:.1f} ms/pass, status={res[1][3]}, kept={res[1][2]}/{T*P}, deterministic={same}"
    if check_gram:
        x = d.features.view(T, -1)
        sub = torch.randperm(T)[:64].to(dev)
        ref = x[sub].double() @ x[sub].double().T
        err = float((sc.G[sub][:, sub] - ref).abs().max() / ref.diagonal().max())
        msg += f", gram err {err:.2e}"
    print(msg, f"(gen {t1 - t0:.0f}s)", flush=True)

for k in (8, 32, 128):                      # config 5: k sweep x token budget {1k, 4k, 16k} at T=600
    for P in (2, 7, 27):
        run(600, P, 1152, k=k, C=48)
run(14400, 16, 1152, C=48)                  # config 4: 2 fps x 2 h
