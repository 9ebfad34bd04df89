"""Pixel-difference kernel overlapped with the k-means solver of the SAME stream: throttled grids x L2 evict-first.
Prints solver / pixel / total device time per combination (CUDA events, 10 reps)."""
import math, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import StreamCompressor, ops, synthetic

T, P, D = 600, 180, 1152
k = math.ceil(T / 15)
dev = torch.device("cuda:0")
st = synthetic.make_stream(T, P, D, seed=0)
d = st.to(dev)
sc = StreamCompressor(T, P, D, st.pixel_values.numel() // (T * P), st.input_ids.numel(), device=dev)
torch.manual_seed(1); random.seed(1)
x2 = d.features.view(T, sc.K)
ops.gram(x2, out=sc.G, workspace=sc.gram_ws)
main = torch.cuda.current_stream()
side = torch.cuda.Stream()

def solve():
    ops.kmeans_solve(sc.G, d.timestamps, sc.km, n_init_given=0, exp_noise=None, reseed_pool=None, rng_state=sc.rng_state)

def run(overlap, reps=10):
    ts = []
    for _ in range(reps + 2):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record(main)
        side.wait_event(e[0])
        if overlap:
            solve(); e[1].record(main)
            e[2].record(side); ops.pixel_diff(d.pixel_values, T, P, sc.W, out=sc.mask, stream=side); e[3].record(side)
            main.wait_event(e[3])
        else:
            solve(); e[1].record(main)
            e[2].record(main); ops.pixel_diff(d.pixel_values, T, P, sc.W, out=sc.mask); e[3].record(main)
        e[4].record(main)
        torch.cuda.synchronize()
        ts.append((e[0].elapsed_time(e[1]) * 1e3, e[2].elapsed_time(e[3]) * 1e3, e[0].elapsed_time(e[4]) * 1e3))
    ts = ts[2:]
    med = lambda i: sorted(t[i] for t in ts)[len(ts) // 2]
    return med(0), med(1), med(2)

print("serial          solver %.1f  pixel %.1f  total %.1f us" % run(False))
for evict in (0, 1):
    for ctas in (0, 296, 148, 96, 64, 48, 32):
        os.environ["CSB_PIX_EVICT"] = str(evict)
        if ctas: os.environ["CSB_PIX_CTAS"] = str(ctas)
        else: os.environ.pop("CSB_PIX_CTAS", None)
        s, p, t = run(True)
        print(f"overlap evict={evict} ctas={ctas or 592:4d}  solver {s:6.1f}  pixel {p:6.1f}  total {t:6.1f} us")
