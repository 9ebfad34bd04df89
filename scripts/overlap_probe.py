"""Where do the kernels of one step sit in time?  CUDA events on both streams relative to a common origin."""
import math, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import StreamCompressor, ops, synthetic

T, P, D = 600, 180, 1152
k = math.ceil(T / 15)
dev = torch.device("cuda:0")
st = synthetic.make_stream(T, P, D, seed=0)
ev, q = synthetic.make_embeddings(k, 3584, seed=0)
d = st.to(dev); ev, q = ev.to(dev), q.to(dev)
sc = StreamCompressor(T, P, D, st.pixel_values.numel() // (T * P), st.input_ids.numel(), device=dev)
torch.manual_seed(1); random.seed(1)

def step(fork):
    sc.phase1(d.features, d.timestamps, pixel_values=d.pixel_values if fork else None)
    sc.phase2(d.features, ev, q, d.pixel_values, d.input_ids, d.attention_mask)

for fork in (False, True, False, True):
    for _ in range(3): step(fork)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step(fork)
    e1.record(); torch.cuda.synchronize()
    print(f"fork={fork}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/step")

# explicit probe: gram, then solver on main + pixel on side, events everywhere
x2 = d.features.view(T, sc.K)
main = torch.cuda.current_stream()
E = {n: torch.cuda.Event(enable_timing=True) for n in ("o", "g1", "s0", "s1", "p0", "p1", "fork")}
for rep in range(3):
    torch.cuda.synchronize()
    E["o"].record(main)
    ops.gram(x2, out=sc.G, workspace=sc.gram_ws)
    E["g1"].record(main)
    sc.km.init_indices[:1].fill_(5); noise = sc.noise[: k - 1]; noise.exponential_(1); sc.pool.random_(0, T)
    E["fork"].record(main)
    E["s0"].record(main)
    ops.kmeans_solve(sc.G, d.timestamps, sc.km, n_init_given=1, exp_noise=noise, reseed_pool=sc.pool)
    E["s1"].record(main)
    sc.side.wait_event(E["fork"])
    E["p0"].record(sc.side)
    ops.pixel_diff(d.pixel_values, T, P, sc.W, out=sc.mask, stream=sc.side)
    E["p1"].record(sc.side)
    torch.cuda.synchronize()
    t = {n: E["o"].elapsed_time(e) * 1e3 for n, e in E.items()}
    print("  ".join(f"{n}={v:7.1f}" for n, v in t.items()))
