"""BASELINE config 4 (T = 14,400 frames, P = 16, D = 1152, k = 960) with the Lloyd loop sharded by frames over the ranks of one box
(SURVEY 8f row 4).  Run under torchrun; every rank builds the same stream and Gram matrix (the per-stream cache is replicated).
Prints one JSON line (rank 0): single-GPU solve vs frame-sharded solve (device time, max over ranks), the time inside the two
all-gathers, and whether the results are bit-identical.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sharded_config4.py
"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from cogstream_b200 import FrameShardedKMeans, ops, synthetic

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
T = int(os.environ.get("T", 14400)); P, D = 16, 1152
k = math.ceil(T / 15)
feats, _ = synthetic.make_features(T, P, D, 0)
ts = (torch.arange(T, dtype=torch.float64) / 2.0).to(dev)
x2 = feats.view(T, -1).to(dev)
G = ops.gram(x2)
state0 = torch.tensor([20261018, 0], dtype=torch.int64, device=dev)

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def timed(fn, reps=3):
    fn(); barrier()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); e0.record(); fn(); e1.record(); barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
        best = ms if best is None else min(best, ms)
    return best

ref = ops.KMeansBuffers(T, k, torch.float64, dev)
def single():
    ops.kmeans_solve(G, ts, ref, n_init_given=0, exp_noise=None, reseed_pool=None, rng_state=state0.clone())
def init_only():
    ops.kmeans_solve(G, ts, ref, n_init_given=0, exp_noise=None, reseed_pool=None, max_iteration=-1, rng_state=state0.clone())
km = FrameShardedKMeans(T, k, torch.float64, dev, rank=rank, world=world)
def sharded():
    km.solve(G, ts, rng_state=state0.clone())

ms_single = timed(single)
ms_init = timed(init_only)
ms_sharded = timed(sharded)
coll = km.collective_ms
if world > 1:
    t = torch.tensor([coll], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); coll = float(t)
single(); sharded(); torch.cuda.synchronize()
a, b = ops.decode_status(ref.status), ops.decode_status(km.buf.status)
same = (torch.equal(ref.assignments, km.buf.assignments) and torch.equal(ref.centroid_times, km.buf.centroid_times)
        and torch.equal(ref.member_offsets, km.buf.member_offsets) and torch.equal(ref.event_order, km.buf.event_order)
        and (a["iterations"], a["n_reseed_used"], a["converged"]) == (b["iterations"], b["n_reseed_used"], b["converged"]))
flag = torch.tensor([1 if same else 0], device=dev)
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps(dict(workload=f"config 4 k-means: T={T} k={k} (Gram replicated, k-means++ replicated, Lloyd sharded by frames)", n_gpus=world,
                          single_gpu_ms=round(ms_single, 3), kmeanspp_only_ms=round(ms_init, 3), frame_sharded_ms=round(ms_sharded, 3),
                          speedup=round(ms_single / ms_sharded, 3), lloyd_single_ms=round(ms_single - ms_init, 3),
                          lloyd_sharded_ms=round(ms_sharded - ms_init, 3), iterations=b["iterations"], reseeds=b["n_reseed_used"],
                          collectives=f"{2 * b['iterations']} x NCCL all_gather_into_tensor (int32 [T] assignments; fp64 [2T + world*k] S entries)",
                          collective_ms_total=round(coll, 3), bit_identical_to_single_gpu=bool(int(flag)))), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
