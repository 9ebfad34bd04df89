"""Copy-only probe for the e2e scaling question: pinned-host -> device copies of one config-2 stream's inputs
(759 MB) per rank, at N = 1/2/4/8 ranks, with and without NUMA-local pinning.  Run under torchrun; rank 0 prints one JSON line
per mode: per-rank GB/s (min / median / max) and the aggregate.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/h2d_probe.py
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
mode = os.environ.get("PROBE_MODE", "default")  # default | numa
node = None
if mode == "numa":
    import bench
    node = bench._set_numa_affinity(local)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NBYTES = 759 * 1000 * 1000
host = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
host.fill_(1)  # first touch after the affinity call
dst = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
back = torch.empty(64 * 1000 * 1000, dtype=torch.uint8).pin_memory()

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

res = {}
for name, fn, nbytes in (("h2d", lambda: dst.copy_(host, non_blocking=True), NBYTES),
                         ("d2h", lambda: back.copy_(dst[: back.numel()], non_blocking=True), back.numel())):
    for _ in range(3):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    gbs = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        vals = sorted(float(v) for v in allv)
    else:
        vals = [gbs]
    res[name] = dict(min=round(vals[0], 2), median=round(vals[len(vals) // 2], 2), max=round(vals[-1], 2), aggregate=round(sum(vals), 1))
info = dict(world=world, mode=mode, numa_node_rank0=node, cpus_rank0=len(os.sched_getaffinity(0)), **res)
if world > 1:
    nodes = [None] * world
    dist.all_gather_object(nodes, (local, node, len(os.sched_getaffinity(0))))
    info["per_rank_node"] = nodes
    dist.barrier(); dist.destroy_process_group()
if rank == 0:
    print(json.dumps(info), flush=True)
