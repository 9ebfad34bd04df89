"""Marginal cost of each stage inside the pipelined (4 streams in flight) steady state: drop one stage at a time."""
import math, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import StreamCompressor, ops, synthetic

T, P, D = 600, 180, 1152
k = math.ceil(T / 15)
dev = torch.device("cuda:0")
st = synthetic.make_stream(T, P, D, seed=0)
ev, q = synthetic.make_embeddings(k, 3584, seed=0)
d = st.to(dev); ev, q = ev.to(dev), q.to(dev)
W = st.pixel_values.numel() // (T * P)
depth = int(os.environ.get("DEPTH", 4))
PRIO = int(os.environ.get("PRIO", 0))   # 1: tensor / latency-bound kernels on a high-priority stream, HBM-bound ones on a low one
lo_streams = {}
lanes = [(torch.cuda.Stream(priority=-1 if PRIO else 0), StreamCompressor(T, P, D, W, st.input_ids.numel(), device=dev)) for _ in range(depth)]
x2 = d.features.view(T, -1)
torch.manual_seed(0); random.seed(0)
for s, sc in lanes:  # one full pass so every buffer holds valid data
    sc.phase1(d.features, d.timestamps); sc.phase2(d.features, ev, q, d.pixel_values, d.input_ids, d.attention_mask)
torch.cuda.synchronize()

class _lo:
    """Run the body on the lane's low-priority stream, ordered after / before the lane's main stream."""
    def __init__(self, sc): self.sc = sc
    def __enter__(self):
        if not PRIO: return
        cur = torch.cuda.current_stream()
        lo = lo_streams.setdefault(id(self.sc), torch.cuda.Stream(priority=0))
        lo.wait_stream(cur)
        self.cur, self.lo = cur, lo
        self.ctx = torch.cuda.stream(lo); self.ctx.__enter__()
    def __exit__(self, *a):
        if not PRIO: return
        self.ctx.__exit__(*a)
        self.cur.wait_stream(self.lo)

def step(sc, skip):
    if "gram" not in skip: ops.gram(x2, out=sc.G, workspace=sc.gram_ws)
    if "solver" not in skip:
        sc.km.init_indices[:1].fill_(7) if not torch.cuda.is_current_stream_capturing() else None; noise = sc.noise[: k - 1]; noise.exponential_(1); sc.pool.random_(0, T)
        ops.kmeans_solve(sc.G, d.timestamps, sc.km, n_init_given=1, exp_noise=noise, reseed_pool=sc.pool)
    if "centroid_gather" not in skip:
      with _lo(sc):
        ops.centroid_gather(x2, sc.km.member_offsets, sc.km.members, sc.km.event_offsets, k, sc.centroids, sc.packed)
    if "small1" not in skip:
        dist = ops.dist_to_assigned_gram(sc.G, sc.km.assignments, sc.km.event_offsets, sc.km.event_order, k)
        sc.reps, sc.rep_counts, sc.rep_flags = ops.select_representatives(dist, sc.km.assignments, k, 2)
    if "small2" not in skip:
        sc.cos, sc.low = ops.cosine_flags(q.view(-1), ev, 0.45)
        sc.minor_flags, sc.minor_list, sc.n_minor = ops.minor_frames(sc.km.assignments, sc.low, sc.rep_flags)
    if "pixel" not in skip:
      with _lo(sc):
        ops.pixel_diff(d.pixel_values, T, P, W, out=sc.mask)
        ops.mask_finalize(sc.mask, T, P, minor_flags=sc.minor_flags)
    if "compact" not in skip:
      with _lo(sc):
        ops.compact_tokens(d.features, T, P, sc.mask, sc.minor_flags, out=sc.out, workspace=sc.compact_ws)
    if "filter" not in skip:
        ops.filter_ids(d.input_ids, sc.image_token_index, sc.mask, d.attention_mask)

USE_GRAPHS = int(os.environ.get("GRAPHS", 1))

def run(skip, n=48):
    main = torch.cuda.current_stream()
    graphs = []
    if USE_GRAPHS:  # one CUDA graph per lane: the host is out of the picture
        for s, sc in lanes:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(s):
                step(sc, skip)  # warm-up outside capture
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                step(sc, skip)
            graphs.append(g)
        torch.cuda.synchronize()
    for w in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for s, _ in lanes: s.wait_event(e0)
        for i in range(n):
            s, sc = lanes[i % depth]
            with torch.cuda.stream(s):
                if USE_GRAPHS: graphs[i % depth].replay()
                else: step(sc, skip)
        for s, _ in lanes:
            ev_ = torch.cuda.Event(); ev_.record(s); main.wait_event(ev_)
        e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

base = run(set())
print(f"depth {depth}: full step {base:.1f} us")
for name in ("gram", "solver", "centroid_gather", "small1", "small2", "pixel", "compact", "filter"):
    t = run({name})
    print(f"  without {name:16s}: {t:7.1f} us  (marginal {base - t:6.1f})")
only = run({"solver", "small1", "small2", "filter"})
print(f"  bandwidth kernels only (gram+centroid_gather+pixel+compact): {only:.1f} us")
allb = {"gram", "centroid_gather", "small1", "small2", "pixel", "compact", "filter"}
print(f"  solver only ({depth} lanes): {run(allb):.1f} us per step")
print(f"  solver + gram only: {run(allb - {'gram'}):.1f} us per step")
print(f"  solver + pixel only: {run(allb - {'pixel'}):.1f} us per step")
print(f"  solver + centroid_gather only: {run(allb - {'centroid_gather'}):.1f} us per step")
