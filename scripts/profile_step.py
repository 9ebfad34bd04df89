"""A few steps of the config-2 pipeline for ncu (launch list / full capture).  Usage: python scripts/profile_step.py [steps]"""
import math
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cogstream_b200 import StreamCompressor, synthetic  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
T, P, D = 600, 180, 1152
k = math.ceil(T / 15)
dev = torch.device("cuda:0")
st = synthetic.make_stream(T, P, D, seed=0)
ev, q = synthetic.make_embeddings(k, 3584, seed=0)
d = st.to(dev)
ev, q = ev.to(dev), q.to(dev)
sc = StreamCompressor(T, P, D, st.pixel_values.numel() // (T * P), st.input_ids.numel(), device=dev)
torch.manual_seed(1234)
random.seed(1234)
for i in range(steps):
    # the pixel-difference mask runs next to the solver (CSB_NOEARLY=1: serial, for a clean solver timeline)
    sc.phase1(d.features, d.timestamps, pixel_values=None if os.environ.get("CSB_NOEARLY") else d.pixel_values)
    sc.phase2(d.features, ev, q, d.pixel_values, d.input_ids, d.attention_mask)
torch.cuda.synchronize()
r = sc.results()
print("ok", r["status"], r["n_kept"], len(r["minor"]))

if os.environ.get("CSB_TIMELINE"):
    names = {1: "start", 2: "init done", 3: "S0 done", 10: "assign", 11: "sync", 12: "rebuild", 13: "S", 14: "sync", 15: "shift", 21: "  pp centre chosen", 22: "  pp row landed, near updated", 23: "  pp partial sums published", 24: "  pp total summed", 25: "  pp own scores done", 26: "  pp warp argmax done", 27: "  pp block sync passed"}
    tl = sc.km.timeline()
    prev = 0
    for tag, t in tl:
        print(f"  {str(names.get(tag, tag)):10s} +{(t - prev) / 1e3:8.2f} us   (t={t / 1e3:8.2f})")
        prev = t
