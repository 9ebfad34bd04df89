"""Does the Gram kernel (148 CTAs x 199 KB shared memory) co-reside with a running k-means solver of another stream?"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import StreamCompressor, ops, synthetic

T, P, D = 600, 180, 1152
k = math.ceil(T / 15)
dev = torch.device("cuda:0")
st = synthetic.make_stream(T, P, D, seed=0)
d = st.to(dev)
W = st.pixel_values.numel() // (T * P)
A = StreamCompressor(T, P, D, W, st.input_ids.numel(), device=dev)
B = StreamCompressor(T, P, D, W, st.input_ids.numel(), device=dev)
x2 = d.features.view(T, -1)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
torch.manual_seed(0)
A.phase1(d.features, d.timestamps); B.phase1(d.features, d.timestamps)
torch.cuda.synchronize()

def solve(sc):
    sc.km.init_indices[:1].fill_(7); noise = sc.noise[: k - 1]; noise.exponential_(1); sc.pool.random_(0, T)
    ops.kmeans_solve(sc.G, d.timestamps, sc.km, n_init_given=1, exp_noise=noise, reseed_pool=sc.pool)

def timed(fn, stream, before=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if before: before()
    with torch.cuda.stream(stream):
        e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3

gram_b = lambda: ops.gram(x2, out=B.G, workspace=B.gram_ws)
pix_b = lambda: ops.pixel_diff(d.pixel_values, T, P, W, out=B.mask)
gath_b = lambda: ops.centroid_gather(x2, B.km.member_offsets, B.km.members, B.km.event_offsets, k, B.centroids, B.packed)
def solver_on_a():
    with torch.cuda.stream(sa):
        solve(A)
for name, fn in (("gram", gram_b), ("pixel_diff", pix_b), ("centroid_gather", gath_b)):
    for _ in range(2): alone = timed(fn, sb)
    for _ in range(2): both = timed(fn, sb, before=solver_on_a)
    print(f"{name:16s} alone {alone:7.1f} us   launched right after a solver started on another stream {both:7.1f} us")
with torch.cuda.stream(sa):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); solve(A); e1.record(); torch.cuda.synchronize()
    print(f"solver alone {e0.elapsed_time(e1)*1e3:.1f} us")
for name, fn in (("gram", gram_b), ("pixel_diff", pix_b), ("centroid_gather", gath_b)):
    # the reverse: the solver's latency while the wide kernel runs (launch wide kernel x3 first, then the solver)
    def wide():
        with torch.cuda.stream(sb):
            for _ in range(4): fn()
    t = timed(lambda: solve(A), sa, before=wide)
    print(f"solver while {name} x4 runs on another stream: {t:7.1f} us")
