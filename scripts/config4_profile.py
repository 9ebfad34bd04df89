"""Stage times of BASELINE config 4 (T=14400 = 2 fps x 2 h, P=16, D=1152, k=960) on one GPU."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import StreamCompressor, ops, synthetic

T, P, D = int(os.environ.get("T", 14400)), 16, 1152
k = math.ceil(T / 15)
dev = torch.device("cuda:0")
st = synthetic.make_stream(T, P, D, 0, C=48)
ev, q = synthetic.make_embeddings(k, 3584, 0)
d = st.to(dev); ev, q = ev.to(dev), q.to(dev)
W = st.pixel_values.numel() // (T * P)
sc = StreamCompressor(T, P, D, W, st.input_ids.numel(), device=dev)
x2 = d.features.view(T, -1)
torch.manual_seed(0)
sc.phase1(d.features, d.timestamps, first_index=5)
sc.phase2(d.features, ev, q, d.pixel_values, d.input_ids, d.attention_mask)
torch.cuda.synchronize()

def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / reps:9.3f} ms", flush=True)

timed("gram", lambda: ops.gram(x2, out=sc.G, workspace=sc.gram_ws))
def solve(max_it):
    ops.kmeans_solve(sc.G, d.timestamps, sc.km, n_init_given=0, exp_noise=None, reseed_pool=None, max_iteration=max_it, rng_state=sc.rng_state)
timed("kmeans (init + 0 iterations)", lambda: solve(0))
timed("kmeans (init + 1 iteration)", lambda: solve(1))
timed("kmeans (full)", lambda: solve(30))
print("status", ops.decode_status(sc.km.status))
timed("centroid_gather", lambda: ops.centroid_gather(x2, sc.km.member_offsets, sc.km.members, sc.km.event_offsets, k, sc.centroids, sc.packed))
timed("phase1 (all)", lambda: sc.phase1(d.features, d.timestamps, first_index=5))
timed("phase2 (all)", lambda: sc.phase2(d.features, ev, q, d.pixel_values, d.input_ids, d.attention_mask))
solve(30); torch.cuda.synchronize()
names = {1: "start", 2: "init done", 3: "S0 done", 10: "assign", 11: "sync", 12: "rebuild", 13: "S", 14: "sync", 15: "shift"}
prev = 0
for tag, t in sc.km.timeline():
    print(f"  {names.get(tag, tag):10s} +{(t - prev) / 1e3:9.2f} us   (t={t / 1e3:9.2f})")
    prev = t
