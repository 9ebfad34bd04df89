#!/usr/bin/env python
"""Benchmark of the Visual Stream Compression hot path (BASELINE.json metric: frames/sec compressed).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2|config3|config4|config5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

The path = phase 1 (Gram -> k-means++ -> Lloyd -> centroids) + phase 2 (relevance -> representatives -> pixel mask ->
pooling -> compaction -> id filtering); the frozen-LLM forwards in between are excluded, their outputs being synthetic inputs.

Workloads (BASELINE.json `configs`):
  config2 (default)  a step = a batch of `--streams-per-step` (32) independent config-2 streams per rank: T=600 frames
                     (1 fps x 10 min), P=180 tokens/frame, D=1152 bf16, k=40, pixels [T*720, 588] bf16, embeddings
                     [40,3584]/[1,3584] bf16, float64 timestamps.  `--depth` (6) streams are in flight per GPU, every lane with
                     its OWN inputs (different seeds: 4.5 GB of distinct inputs per GPU), and the step ends with the one
                     collective of the path, the all-gather of the per-stream stats rows, inside the timed region.
  config3            256 concurrent config-2 streams sharded over the ranks through the public StreamPool API (strong scaling)
  config4            one long stream per rank (T=14,400 = 2 fps x 2 h, P=16, k=960); a step = one new question: append the last
                     `--segment` frames to the per-stream cache (incremental Gram), re-cluster everything, compress
  config5            sweep k in {8,32,128} x token budget {1k,4k,16k} (T=600); a step = the 9 points once

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = through the public pipeline API with pinned HOST
inputs: H2D copies of every input and D2H of the results (stats, token mask, compressed features, filtered ids) inside the
timed region.  `--impl reference` times the CPU restatement of the reference (oracle/, kind "port": the reference is Python and
/root/reference is not on the GPU box) on the host cores, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "frames/sec compressed"
UNIT = "frames/s"
CFG2 = dict(T=600, P=180, D=1152, merge=2, C=588, Dl=3584, fps=1.0)
CFG4 = dict(T=14400, P=16, D=1152, merge=2, C=588, Dl=3584, fps=2.0)
SWEEP = [(k, P) for k in (8, 32, 128) for P in (2, 7, 27)]  # token budget T*P ~ 1k / 4k / 16k at T=600


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]),
                    bf16_tflops_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                    power_w_max=max(power) if power else None)


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, all host threads
# ----------------------------------------------------------------------------------------------------------
def _cpu_path_once(st, ev, q, k, seed, *, cluster_num=None, init_indices=None):
    from cogstream_b200 import synthetic
    from oracle import cogstream_oracle as orc

    random.seed(seed)
    torch.manual_seed(seed)
    t0 = time.perf_counter()
    kw = {} if init_indices is None else dict(init_indices=init_indices)
    p1 = orc.phase1(st.features, st.T, st.timestamps, cluster_num=cluster_num, **kw)  # k-means++ sampled like the reference does
    p2 = orc.phase2(st.features, p1, ev, q, st.pixel_values, st.grid_sizes, st.merge_sizes, st.modals, st.input_ids,
                    synthetic.IMAGE_TOKEN_INDEX, st.attention_mask)
    dt = time.perf_counter() - t0
    return dt, p1, p2


def _make_cpu_sample(w, P_sample, seed=0):
    from cogstream_b200 import synthetic

    st = synthetic.make_stream(w["T"], P_sample, w["D"], seed, fps=w["fps"], merge=w["merge"], C=w["C"])
    k = math.ceil(w["T"] / 15)
    ev, q = synthetic.make_embeddings(k, w["Dl"], seed)
    return st, ev, q, k


def _cpu_config2(budget_s: float):
    """One config-2 stream on the CPU, P reduced until a run fits the budget (the CPU cost is ~linear in P)."""
    w = CFG2
    st, ev, q, k = _make_cpu_sample(w, 6)
    t_small = min(_cpu_path_once(st, ev, q, k, 0)[0], _cpu_path_once(st, ev, q, k, 0)[0])
    P_s = int(max(6, min(w["P"], budget_s / max(t_small / 6.0, 1e-6))))
    if P_s >= w["P"] * 0.9:
        P_s = w["P"]
    st, ev, q, k = _make_cpu_sample(w, P_s)

    def run(seed):
        return _cpu_path_once(st, ev, q, k, seed)[0]

    sample = (f"1 stream of config 2 with T={w['T']} k={k} D={w['D']} and P={P_s} of {w['P']} tokens/frame (CPU cost ~linear in P; "
              f"value scaled by P_s/P to the full workload when P_s < P), k-means++ sampled")
    return run, w["T"], P_s / w["P"], sample


def _cpu_config4(budget_s: float):
    """Config 4 on the CPU.  The reference's own k-means++ is O(k^2 T K) (it re-computes the distances to ALL chosen centres every
    round, kmeans_with_time.py:46-48): 2.4e17 flop at T=14,400, k=960, K=18,432 -- days of CPU time.  The bounded sample is the
    FULL-size Lloyd loop + phase 2 with the init injected, i.e. the CPU number EXCLUDES the init and overstates the reference."""
    from cogstream_b200 import synthetic
    w = CFG4
    st = synthetic.make_stream(w["T"], w["P"], w["D"], 0, fps=w["fps"], merge=w["merge"], C=w["C"])
    k = math.ceil(w["T"] / 15)
    ev, q = synthetic.make_embeddings(k, w["Dl"], 0)
    g = torch.Generator().manual_seed(0)
    init = torch.randperm(w["T"], generator=g)[:k].tolist()

    def run(seed):
        return _cpu_path_once(st, ev, q, k, seed, init_indices=init)[0]

    sample = (f"1 stream of config 4 at full size (T={w['T']} P={w['P']} D={w['D']} k={k}), Lloyd loop + selection + phase 2 with the "
              "init INJECTED (the reference's O(k^2 T K) k-means++ = 2.4e17 flop is not timeable): the CPU value excludes the init")
    return run, w["T"], 1.0, sample


def _cpu_config5(budget_s: float):
    from cogstream_b200 import synthetic
    pts = []
    for kk, P in SWEEP:
        st = synthetic.make_stream(600, P, 1152, seed=kk * 100 + P)
        ev, q = synthetic.make_embeddings(kk, 3584, seed=kk * 100 + P)
        pts.append((st, ev, q, kk))

    def run(seed):
        return sum(_cpu_path_once(st, ev, q, kk, seed, cluster_num=kk)[0] for st, ev, q, kk in pts)

    return run, 600 * len(SWEEP), 1.0, "the 9 sweep points (k x token budget) once, full size, k-means++ sampled"


_CPU_WORKLOADS = {"config2": _cpu_config2, "config3": _cpu_config2, "config4": _cpu_config4, "config5": _cpu_config5}


def cpu_baseline(workload: str, budget_s: float):
    torch.set_num_threads(os.cpu_count() or 1)
    run, frames, scale, sample = _CPU_WORKLOADS[workload](budget_s)
    t = run(0)
    return dict(value=frames / t * scale, unit=UNIT, cores=torch.get_num_threads(), kind="port", sample=sample + "; 1 run",
                seconds_per_run=t, measured_sample_frames_per_s=frames / t)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    per_step = 150.0 / max(args.steps + args.warmup, 1)
    run, frames, scale, sample = _CPU_WORKLOADS[args.workload](per_step)
    for i in range(args.warmup if args.workload != "config4" else min(args.warmup, 1)):
        run(100 + i)
    steps = args.steps if args.workload != "config4" else min(args.steps, 3)
    t0 = time.perf_counter()
    for i in range(steps):
        run(i)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    value = frames / dt * scale
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=args.warmup, ms_per_step=dt * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic", impl="reference",
                config=_config_dict(args),
                cpu_baseline=dict(value=value, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                                  sample="each step = " + sample, measured_sample_frames_per_s=frames / dt),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    _emit(line)


def _config_dict(args, world=None):
    wl = args.workload
    if wl in ("config2", "config3"):
        w = CFG2
        d = dict(T=w["T"], P=w["P"], D=w["D"], k=math.ceil(w["T"] / 15), pixel_row=f"[T*{w['P'] * w['merge'] ** 2}, {w['C']}] bf16",
                 embed_dim=w["Dl"], timestamps="float64", init="k-means++ sampled on device by the solver's built-in Philox generator (first centre, Exp(1) noise, re-seeds)",
                 l2=f"every lane has its own inputs (757 MB per stream, {args.depth} lanes = {0.757 * args.depth:.1f} GB distinct per GPU) "
                    "> 126 MB L2; no explicit flush")
        if wl == "config2":
            d.update(workload=f"BASELINE config 2: independent streams of 1 fps x 10 min; a step = {args.streams_per_step} streams per GPU, "
                              f"{args.depth} in flight, stats all-gather at the end of every step",
                     streams_per_step=args.streams_per_step, parallelism="independent streams, no data-path collective")
        else:
            d.update(workload="BASELINE config 3: 256 concurrent config-2 streams sharded over the ranks (StreamPool.run_batch), one "
                              "stats all-gather per step", streams_per_step_total=256, parallelism="stream s -> rank s mod W")
        return d
    if wl == "config4":
        w = CFG4
        return dict(workload=f"BASELINE config 4: one long stream per GPU, 2 fps x 2 h; a step = one new question: the last {args.segment} "
                             "frames are appended to the per-stream cache (incremental Gram), all frames are re-clustered (fresh k-means++), "
                             "compressed", T=w["T"], P=w["P"], D=w["D"], k=math.ceil(w["T"] / 15), segment=args.segment,
                    timestamps="float64", l2="X 531 MB, G 1.66 GB, pixels 1.08 GB > L2", parallelism="replicas (one stream per GPU)")
    return dict(workload="BASELINE config 5: sweep k in {8,32,128} x token budget {1k,4k,16k} at T=600 (P in {2,7,27}); a step = the 9 "
                         "points once, one stream each, serial on one CUDA stream", points=[dict(k=k, P=P) for k, P in SWEEP],
                D=1152, l2="flushed between steps is unnecessary: latency-bound sizes; inputs 9 x (2-30 MB)", parallelism="replicas")


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch.distributed as dist
        from cogstream_b200 import _lib
        self.args = args
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; cogstream_b200 has no CPU fallback (use --impl reference for the CPU arm)")
        _set_numa_affinity(self.local_rank)
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner / warnings never go to stdout
            dist.init_process_group("nccl", device_id=self.dev)
        self.lib = _lib.load()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed(self, step_fn, steps, warmup, join_fn=None, streams=()):
        """W untimed + K timed steps.  step_fn(i) enqueues one step (and may block the host); join_fn() makes the current stream wait
        for everything the steps enqueued on the other `streams` (which in turn wait for the start event).  Barrier + synchronize
        on both sides; CUDA events on the current stream; max over ranks.  Returns (ms for the K steps, launches of our kernels
        counted by the library)."""
        main = torch.cuda.current_stream(self.dev)
        for i in range(warmup):
            step_fn(i)
        if join_fn:
            join_fn()
        self.barrier()
        l0 = self.lib.csb_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for st in streams:
            st.wait_event(e0)
        for i in range(steps):
            step_fn(warmup + i)
        if join_fn:
            join_fn()
        e1.record(main)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)), self.lib.csb_kernel_launches() - l0


def _set_numa_affinity(local_rank: int):
    """Pin this rank's host threads to the CPUs of the NUMA node its GPU hangs off BEFORE any pinned allocation, so that the
    first-touch policy places the pinned staging buffers in the memory next to the GPU's PCIe root (e2e H2D copies)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        node = None
        try:
            node = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            path = f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node"
            if os.path.isfile(path):
                node = int(open(path).read().strip())
        if node is None or node < 0:
            return None
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return node
    except Exception:
        return None


class Lane:
    """One in-flight stream slot: its own CUDA stream, inputs (device-resident + pinned host copies, seed = 100 * rank + lane),
    workspaces (StreamCompressor) and, for e2e, device staging + pinned result buffers."""

    def __init__(self, ctx: Ctx, w, index: int, cfg=None, P=None, make_host=True):
        from cogstream_b200 import StreamCompressor, synthetic
        dev = ctx.dev
        self.ctx = ctx
        T, P, D = w["T"], P or w["P"], w["D"]
        self.T, self.P, self.D = T, P, D
        seed = 100 * ctx.rank + index
        st = synthetic.make_stream(T, P, D, seed=seed, fps=w["fps"], merge=w["merge"], C=w["C"])
        self.k = cfg.clusters_for(T) if cfg is not None else math.ceil(T / 15)
        ev, q = synthetic.make_embeddings(self.k, w["Dl"], seed=seed)
        self.W = st.pixel_values.numel() // (T * P)
        self.n_ids = st.input_ids.numel()
        self.stream = torch.cuda.Stream(device=dev)
        self.d = st.to(dev)
        self.ev_d, self.q_d = ev.to(dev), q.to(dev)
        self.host = None
        if make_host:
            self.host = st.pin()
            self.ev_h, self.q_h = ev.pin_memory(), q.pin_memory()
        self.sc = StreamCompressor(T, P, D, self.W, self.n_ids, device=dev, cfg=cfg)
        self.staged = False
        self.pending = False
        self.graphs = False

    def capture(self):
        with torch.cuda.stream(self.stream):
            d = self.d
            self.sc.capture(d.features, d.timestamps, self.ev_d, self.q_d, d.pixel_values, d.input_ids, d.attention_mask,
                            early_pixel=bool(self.ctx.args.early_pixel))
        self.graphs = True

    def run_resident(self):
        if self.graphs:
            self.sc.replay1()
            self.sc.replay2()
            return self.sc._stats_static
        d = self.d
        self.sc.phase1(d.features, d.timestamps, pixel_values=d.pixel_values if self.ctx.args.early_pixel else None)
        self.sc.phase2(d.features, self.ev_d, self.q_d, d.pixel_values, d.input_ids, d.attention_mask)
        return self.sc.stats_row()

    # ---- e2e: pinned host -> device staging -> pipeline -> pinned host --------------------------------------
    def stage(self):
        if self.staged:
            return
        from cogstream_b200 import STATS_FIELDS
        d = self.d
        self.e_feat = torch.empty_like(d.features); self.e_pix = torch.empty_like(d.pixel_values)
        self.e_ts = torch.empty_like(d.timestamps); self.e_ids = torch.empty_like(d.input_ids)
        self.e_am = torch.empty_like(d.attention_mask); self.e_ev = torch.empty_like(self.ev_d); self.e_q = torch.empty_like(self.q_d)
        n = self.T * self.P
        self.r_stats = torch.empty(len(STATS_FIELDS), dtype=torch.int64).pin_memory()
        self.r_mask = torch.empty(n, dtype=torch.uint8).pin_memory()
        self.r_feat = torch.empty(n, self.D, dtype=torch.bfloat16).pin_memory()  # worst case; the copy is n_kept rows
        self.r_ids = torch.empty(self.n_ids, dtype=torch.int64).pin_memory()
        self.r_am = torch.empty(self.n_ids, dtype=torch.int64).pin_memory()
        self.staged = True

    def h2d_bytes(self):
        h = self.host
        return sum(t.numel() * t.element_size() for t in (h.features, h.pixel_values, h.timestamps, h.input_ids, h.attention_mask,
                                                           self.ev_h, self.q_h))

    def e2e_enqueue(self):
        """Everything up to the stats row; the caller later calls e2e_finish() (host wait + exact-size result copies)."""
        self.stage()
        h = self.host
        self.e_feat.copy_(h.features, non_blocking=True)
        self.e_ts.copy_(h.timestamps, non_blocking=True)
        self.e_pix.copy_(h.pixel_values, non_blocking=True)
        self.sc.phase1(self.e_feat, self.e_ts, pixel_values=self.e_pix if self.ctx.args.early_pixel else None)
        self.e_ids.copy_(h.input_ids, non_blocking=True)
        self.e_am.copy_(h.attention_mask, non_blocking=True)
        self.e_ev.copy_(self.ev_h, non_blocking=True)
        self.e_q.copy_(self.q_h, non_blocking=True)
        self.sc.phase2(self.e_feat, self.e_ev, self.e_q, self.e_pix, self.e_ids, self.e_am)
        self.r_stats.copy_(self.sc.stats_row(), non_blocking=True)
        self.pending = True

    def e2e_finish(self):
        """What a caller of the public API does with the result: wait for the stats row (n_kept, n_ids_out), then read the
        compressed features, the filtered ids / attention mask and the token mask back.  Returns the D2H bytes."""
        from cogstream_b200 import STATS_FIELDS
        if not self.pending:
            return 0
        self.stream.synchronize()
        nk = int(self.r_stats[STATS_FIELDS.index("n_kept")])
        no = int(self.r_stats[STATS_FIELDS.index("n_ids_out")])
        with torch.cuda.stream(self.stream):
            self.r_feat[:nk].copy_(self.sc.out[:nk], non_blocking=True)
            self.r_ids[:no].copy_(self.sc.out_ids[:no], non_blocking=True)
            self.r_am[:no].copy_(self.sc.out_attention[:no], non_blocking=True)
            self.r_mask.copy_(self.sc.mask, non_blocking=True)
        self.pending = False
        return self.r_stats.numel() * 8 + nk * self.D * 2 + 2 * no * 8 + self.r_mask.numel()


def run_streams_workload(ctx: Ctx, args):
    """config2 / config3: independent streams of the config-2 shape."""
    from cogstream_b200 import STATS_FIELDS, StreamInputs, StreamPool, gather_stats
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    w = CFG2
    T = w["T"]
    cfg3 = args.workload == "config3"
    S = (256 + world - 1) // world if cfg3 else args.streams_per_step
    frames_per_step_total = 256 * T if cfg3 else world * S * T
    depth = max(1, args.depth)
    lanes = [Lane(ctx, w, i) for i in range(depth)]
    torch.manual_seed(1234 + rank)
    random.seed(1234 + rank)
    F = len(STATS_FIELDS)
    main = torch.cuda.current_stream(dev)

    use_graphs = False
    launches_per_stream = None
    if args.graphs and not cfg3:
        try:
            c0 = ctx.lib.csb_kernel_launches()
            for L in lanes:
                L.capture()
            torch.cuda.synchronize()
            launches_per_stream = (ctx.lib.csb_kernel_launches() - c0) // (2 * len(lanes))  # one eager + one captured pass per lane
            use_graphs = True
        except Exception as exc:  # capture is an optimisation of the launch path only
            print(f"bench.py: CUDA graph capture unavailable ({type(exc).__name__}: {exc}); running eager", file=sys.stderr)
            torch.cuda.synchronize()
            for L in lanes:
                L.graphs = False

    rows_buf = [torch.zeros(S, F, dtype=torch.int64, device=dev) for _ in range(2)]
    gathered = [None]
    g_done = [None, None]

    def join(use):
        for L in use:
            done = torch.cuda.Event()
            done.record(L.stream)
            main.wait_event(done)

    def make_step(use, gather=True):
        counter = [0]

        def step(s):
            rows = rows_buf[s & 1]
            for i in range(S):
                L = use[counter[0] % len(use)]
                counter[0] += 1
                with torch.cuda.stream(L.stream):
                    if i < len(use) and g_done[s & 1] is not None:
                        L.stream.wait_event(g_done[s & 1])  # rows_buf[s & 1] was last read by the gather of step s - 2
                    rows[i].copy_(L.run_resident(), non_blocking=True)
            if gather:
                join(use)
                gathered[0] = gather_stats(rows)  # the one collective of the path (NCCL all-gather over NVLink), once per step
                ev = torch.cuda.Event()
                ev.record(main)
                g_done[s & 1] = ev
        return step

    def timed_lanes(step, steps, warmup, use):
        return ctx.timed(step, steps, warmup, join_fn=lambda: join(use), streams=[L.stream for L in use])

    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    W_ = max(args.warmup, 3)

    if cfg3:
        pool = StreamPool(T, w["P"], w["D"], lanes[0].W, lanes[0].n_ids, device=dev, depth=depth)
        ins = []
        embeds = []
        for i in range(S):
            L = lanes[i % depth]
            ins.append(StreamInputs(L.d.features, L.d.timestamps, L.d.pixel_values, L.d.input_ids, L.d.attention_mask))
            embeds.append((L.ev_d, L.q_d))

        def step3(s):
            it = iter(embeds)
            _, rows = pool.run_batch(ins, lambda sc, x: next(it), keep_outputs=False)
            gathered[0] = gather_stats(rows)

        ms_total, launches = ctx.timed(step3, args.steps, W_)
        stat_lane = pool.lanes[0].sc
    else:
        ms_total, launches = timed_lanes(make_step(lanes), args.steps, W_, lanes)
        if use_graphs:
            launches = launches_per_stream * S * args.steps  # graph replays re-launch the captured kernels without the host-side counter
        stat_lane = lanes[0].sc
    stats_row0 = stat_lane.stats_row().cpu()
    g_done[0] = g_done[1] = None
    ms_single, _ = timed_lanes(make_step(lanes[:1], gather=False), max(2, args.steps // 4), 1, lanes[:1])
    ms_single /= max(2, args.steps // 4) * S  # per stream, nothing overlapped

    # ---- informational: the same run without materialising the fp32 centroids.  The reference returns them
    # (kmeans_with_time.py:134) and its only consumer on this path is select_additional_frames (cogreasoner_chat.py:50-64); the
    # fused pipeline picks the representatives from the Gram matrix, so a caller that does not look at the centroids can skip the
    # 282 MB pass (`StreamConfig.compute_centroids=False`).  NOT the headline: `value` above always includes the pass.
    lean = None
    if use_graphs and not cfg3:
        from cogstream_b200 import StreamCompressor, StreamConfig
        saved = [L.sc for L in lanes]
        try:
            for L in lanes:
                L.sc = StreamCompressor(L.T, L.P, L.D, L.W, L.n_ids, device=dev, cfg=StreamConfig(compute_centroids=False))
                L.capture()
            torch.cuda.synchronize()
            g_done[0] = g_done[1] = None
            n_lean = max(3, args.steps // 2)
            ms_lean, _ = timed_lanes(make_step(lanes), n_lean, 2, lanes)
            lean = dict(value=frames_per_step_total / (ms_lean / n_lean * 1e-3), unit=UNIT, ms_per_stream=ms_lean / (n_lean * S), steps=n_lean,
                        note="informational, not the headline: StreamConfig(compute_centroids=False) -- the fp32 centroids the reference "
                             "returns are not materialised (nothing downstream on the fused path reads them: representatives come from "
                             "the Gram matrix); every other output is identical")
        finally:
            for L, sc in zip(lanes, saved):
                L.sc = sc
            g_done[0] = g_done[1] = None
        del saved

    # ---- e2e: pinned host inputs, H2D + pipeline + D2H of the results, through the public API ----
    d2h_bytes = [0]

    def step_e2e(s):
        total = 0
        for i in range(S):
            L = lanes[i % depth]
            total += L.e2e_finish()
            with torch.cuda.stream(L.stream):
                L.e2e_enqueue()
        rows = rows_buf[s & 1]
        for i, L in enumerate(lanes):
            total += L.e2e_finish()
        join(lanes)
        for i in range(S):
            rows[i].copy_(lanes[i % depth].r_stats, non_blocking=True)  # (the last stream of each lane; shape of the collective is what matters)
        gathered[0] = gather_stats(rows)
        d2h_bytes[0] = total

    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    ms_e2e, _ = ctx.timed(step_e2e, e2e_steps, 1, join_fn=lambda: join(lanes), streams=[L.stream for L in lanes])
    clocks = sampler.stop() if sampler else None
    torch.cuda.synchronize()

    ms_step = ms_total / args.steps
    value = frames_per_step_total / (ms_step * 1e-3)
    ms_e2e_step = ms_e2e / e2e_steps
    e2e_value = frames_per_step_total / (ms_e2e_step * 1e-3)
    h2d = lanes[0].h2d_bytes() * S

    stages = roofline = cpu = None
    if rank == 0:
        L0 = lanes[0]
        stages = _stage_times(L0.sc, L0.d, L0.ev_d, L0.q_d, reps=max(3, min(args.steps, 10)))
        roofline = _roofline(stages, L0.sc, ms_stream=ms_step / S, ms_single=ms_single)
        roofline["neighbour_ops"] = _neighbour_ops(L0.sc, L0.d)
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(args.workload, budget_s=20.0)
    _finish(ctx)
    if rank != 0:
        return
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=W_, ms_per_step=ms_step,
                higher_is_better=True, scaling="strong" if cfg3 else "weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=_config_dict(args), clocks=clocks,
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=int(d2h_bytes[0]), ms_per_step=ms_e2e_step,
                         steps=e2e_steps, note="pinned host inputs (NUMA-local to the GPU) -> H2D -> phase 1 + 2 -> host wait on the stats row -> "
                                               "D2H of compressed features, filtered ids / attention mask, token mask"),
                gpu_launches=int(launches), roofline=roofline, stages_ms=stages,
                single_stream=dict(ms=ms_single, value=T / (ms_single * 1e-3), unit=UNIT,
                                   hbm_frac=roofline["whole_path"]["single_stream"]["frac"] if roofline else None,
                                   note="one stream at a time on one CUDA stream (BASELINE config 2 read literally): latency of the path"),
                pipeline=dict(depth=len(lanes), cuda_graphs=use_graphs, early_pixel_diff=bool(args.early_pixel), streams_per_step=S, ms_per_stream=ms_step / S,
                              timed_region_ms=ms_total, launches_per_stream=launches_per_stream,
                              note="successive streams are independent; `depth` of them are in flight on separate CUDA streams with their own "
                                   "inputs and workspaces; no synchronisation between steps other than the stats all-gather"),
                stream_stats=dict(zip(STATS_FIELDS, [int(v) for v in stats_row0.tolist()])),
                all_gathered_stats_shape=list(gathered[0].shape) if gathered[0] is not None else None)
    if lean is not None:
        line["without_centroid_pass"] = lean
    if cpu is not None:
        line["cpu_baseline"] = cpu
    _emit(line)


def _finish(ctx):
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------
# config 4: one long stream, incremental re-clustering per question
# ----------------------------------------------------------------------------------------------------------
def run_config4(ctx: Ctx, args):
    from cogstream_b200 import STATS_FIELDS, StreamCache, StreamCompressor, StreamConfig, gather_stats, ops, synthetic
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    w = CFG4
    T, P, D = w["T"], w["P"], w["D"]
    k = math.ceil(T / 15)
    seg = args.segment
    st = synthetic.make_stream(T, P, D, seed=rank, fps=w["fps"], merge=w["merge"], C=w["C"])
    ev, q = synthetic.make_embeddings(k, w["Dl"], seed=rank)
    host = st.pin()
    d = st.to(dev)
    ev_d, q_d = ev.to(dev), q.to(dev)
    Wp = st.pixel_values.numel() // (T * P)
    torch.manual_seed(1234 + rank)
    random.seed(1234 + rank)
    cache = StreamCache(T, P, D, device=dev)
    cache.append(d.features[: (T - seg) * P], d.timestamps[: T - seg])  # history up to the previous question
    sc = StreamCompressor(T, P, D, Wp, st.input_ids.numel(), device=dev, cfg=StreamConfig(compute_centroids=True))
    x_all = cache.X[:T]
    gathered = [None]

    def question(incremental=True, feats_new=None, ts_new=None):
        """A new question arrives at frame T: the frames since the last one are appended, everything is re-clustered."""
        if incremental:
            cache.T = T - seg
            cache.append(d.features[(T - seg) * P:] if feats_new is None else feats_new, d.timestamps[T - seg:] if ts_new is None else ts_new)
            G = cache.gram
        else:
            G = None
        sc.phase1(x_all.view(T * P, D), cache.ts[:T], first_index="device", gram=G)
        sc.phase2(x_all.view(T * P, D), ev_d, q_d, d.pixel_values, d.input_ids, d.attention_mask)

    def step_inc(s):
        question(True)
        gathered[0] = gather_stats(sc.stats_row().unsqueeze(0))

    def step_scratch(s):
        question(False)
        gathered[0] = gather_stats(sc.stats_row().unsqueeze(0))

    # e2e: only the new segment crosses PCIe (features, timestamps, pixels of the new frames), results come back
    G4 = P * w["merge"] ** 2
    e_feat = torch.empty(seg * P, D, dtype=torch.bfloat16, device=dev)
    e_ts = torch.empty(seg, dtype=torch.float64, device=dev)
    r_stats = torch.empty(len(STATS_FIELDS), dtype=torch.int64).pin_memory()
    r_feat = torch.empty(T * P, D, dtype=torch.bfloat16).pin_memory()
    r_ids = torch.empty(st.input_ids.numel(), dtype=torch.int64).pin_memory()
    r_mask = torch.empty(T * P, dtype=torch.uint8).pin_memory()
    e_ev, e_q = torch.empty_like(ev_d), torch.empty_like(q_d)
    ev_h, q_h = ev.pin_memory(), q.pin_memory()
    h2d = (seg * P * D * 2 + seg * 8 + seg * G4 * w["C"] * 2 + ev_h.numel() * 2 + q_h.numel() * 2)
    d2h = [0]

    def step_e2e(s):
        e_feat.copy_(host.features[(T - seg) * P:], non_blocking=True)
        e_ts.copy_(host.timestamps[T - seg:], non_blocking=True)
        d.pixel_values[(T - seg) * G4:].copy_(host.pixel_values[(T - seg) * G4:], non_blocking=True)
        e_ev.copy_(ev_h, non_blocking=True); e_q.copy_(q_h, non_blocking=True)
        question(True, e_feat, e_ts)
        r_stats.copy_(sc.stats_row(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        nk, no = int(r_stats[STATS_FIELDS.index("n_kept")]), int(r_stats[STATS_FIELDS.index("n_ids_out")])
        r_feat[:nk].copy_(sc.out[:nk], non_blocking=True)
        r_ids[:no].copy_(sc.out_ids[:no], non_blocking=True)
        r_mask.copy_(sc.mask, non_blocking=True)
        gathered[0] = gather_stats(sc.stats_row().unsqueeze(0))
        d2h[0] = r_stats.numel() * 8 + nk * D * 2 + no * 8 + r_mask.numel()

    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    W_ = max(args.warmup, 3)
    ms_inc, launches = ctx.timed(step_inc, args.steps, W_)
    row0 = sc.stats_row().cpu()
    ms_scr, _ = ctx.timed(step_scratch, max(2, args.steps // 2), 1)
    ms_e2e, _ = ctx.timed(step_e2e, max(2, args.steps // 2), 1)
    clocks = sampler.stop() if sampler else None
    ms_step = ms_inc / args.steps
    ms_scr /= max(2, args.steps // 2)
    ms_e2e /= max(2, args.steps // 2)

    stages = roofline = cpu = None
    if rank == 0:
        stages = _stage_times_config4(sc, cache, d, ev_d, q_d, seg)
        pk = _peaks()
        K = P * D
        gram_full_flops = T * (T + 1) * K
        dom = max(stages, key=lambda n: stages[n])
        gram_tf = gram_full_flops / (stages["gram_full"] * 1e-3) / 1e12
        roofline = dict(bound="tensor", kernel="gram_tc_kernel (long-stream direct path)", achieved=gram_tf, peak=pk["bf16_tflops"], unit="TFLOP/s",
                        frac=gram_tf / pk["bf16_tflops"], traffic=None, peaks=pk["source"], dominant_stage=dom,
                        frac_of_sustained=gram_tf / pk["bf16_tflops_sustained"], algorithmic_flops=gram_full_flops,
                        note="roofline of the tensor-bound kernel of this workload (from-scratch Gram, k=960 -> SURVEY 8d: tensor-bound); the "
                             "k-means++ / Lloyd kernels are latency-bound chains of 959 sampling rounds / grid barriers")
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline("config4", budget_s=30.0)
    _finish(ctx)
    if rank != 0:
        return
    line = dict(metric=METRIC, value=world * T / (ms_step * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps, warmup=W_, ms_per_step=ms_step,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic", config=_config_dict(args), clocks=clocks,
                e2e=dict(value=world * T / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h[0]), ms_per_step=ms_e2e,
                         note="only the new segment (features, timestamps, pixels) crosses PCIe; the history lives in the per-stream device cache"),
                gpu_launches=int(launches), roofline=roofline, stages_ms=stages,
                config4=dict(incremental_ms=ms_step, from_scratch_ms=ms_scr, new_frames_per_question=seg,
                             from_scratch_frames_per_s=world * T / (ms_scr * 1e-3)),
                stream_stats=dict(zip(STATS_FIELDS, [int(v) for v in row0.tolist()])),
                all_gathered_stats_shape=list(gathered[0].shape))
    if cpu is not None:
        line["cpu_baseline"] = cpu
    _emit(line)


def _ev_time(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 5)


def _stage_times_config4(sc, cache, d, ev_d, q_d, seg, reps=3):
    from cogstream_b200 import ops
    T, P, k = sc.T, sc.P, sc.k
    x2 = cache.X[:T]

    def solve(max_it):
        ops.kmeans_solve(cache.gram, cache.ts[:T], sc.km, n_init_given=0, exp_noise=None, reseed_pool=None, max_iteration=max_it,
                         rng_state=sc.rng_state)

    def ext():
        cache.T = T - seg
        cache.append(d.features[(T - seg) * P:], d.timestamps[T - seg:])

    out = dict(gram_full=_ev_time(lambda: ops.gram(x2, out=sc.G, workspace=sc.gram_ws), reps),
               gram_extend=_ev_time(ext, reps),
               kmeanspp=_ev_time(lambda: solve(0), reps))
    out["kmeans_full"] = _ev_time(lambda: solve(30), reps)
    out["lloyd"] = round(out["kmeans_full"] - out["kmeanspp"], 5)
    out["phase2"] = _ev_time(lambda: sc.phase2(x2.view(T * P, sc.D), ev_d, q_d, d.pixel_values, d.input_ids, d.attention_mask), reps)
    return out


# ----------------------------------------------------------------------------------------------------------
# config 5: sweep
# ----------------------------------------------------------------------------------------------------------
def run_config5(ctx: Ctx, args):
    from cogstream_b200 import STATS_FIELDS, StreamConfig, gather_stats
    rank, world = ctx.rank, ctx.world
    lanes = []
    for kk, P in SWEEP:
        L = Lane(ctx, CFG2, kk * 100 + P, cfg=StreamConfig(cluster_num=kk), P=P, make_host=True)
        L.stream = torch.cuda.current_stream(ctx.dev)
        lanes.append(L)
    torch.manual_seed(1234 + rank)
    random.seed(1234 + rank)
    use_graphs = False
    launches_per_step = None
    if args.graphs:
        try:
            c0 = ctx.lib.csb_kernel_launches()
            for L in lanes:
                d = L.d
                L.sc.capture(d.features, d.timestamps, L.ev_d, L.q_d, d.pixel_values, d.input_ids, d.attention_mask,
                             early_pixel=bool(args.early_pixel))
                L.graphs = True
            launches_per_step = (ctx.lib.csb_kernel_launches() - c0) // 2  # one eager + one captured pass per point
            use_graphs = True
        except Exception as exc:
            print(f"bench.py: CUDA graph capture unavailable ({type(exc).__name__}: {exc}); running eager", file=sys.stderr)
            for L in lanes:
                L.graphs = False
    rows = torch.zeros(len(lanes), len(STATS_FIELDS), dtype=torch.int64, device=ctx.dev)
    gathered = [None]

    def step(s):
        for i, L in enumerate(lanes):
            rows[i].copy_(L.run_resident(), non_blocking=True)
        gathered[0] = gather_stats(rows)

    d2h = [0]

    def step_e2e(s):
        tot = 0
        for L in lanes:
            L.e2e_enqueue()
            tot += L.e2e_finish()
        gathered[0] = gather_stats(rows)
        d2h[0] = tot

    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    W_ = max(args.warmup, 3)
    ms, launches = ctx.timed(step, args.steps, W_)
    if use_graphs:
        launches = launches_per_step * args.steps  # graph replays re-launch the captured kernels without the host-side counter
    ms_e2e, _ = ctx.timed(step_e2e, args.steps, 1)
    clocks = sampler.stop() if sampler else None
    pts = None
    if rank == 0:
        pts = []
        for (kk, P), L in zip(SWEEP, lanes):
            t = _ev_time(L.run_resident, 10)
            pts.append(dict(k=kk, P=P, token_budget=600 * P, gpu_ms=t, gpu_frames_per_s=600 / (t * 1e-3)))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from cogstream_b200 import synthetic
        torch.set_num_threads(os.cpu_count() or 1)
        tot = 0.0
        for (kk, P), pt in zip(SWEEP, pts):
            st = synthetic.make_stream(600, P, 1152, seed=kk * 100 + P)
            ev, q = synthetic.make_embeddings(kk, 3584, seed=kk * 100 + P)
            dt = _cpu_path_once(st, ev, q, kk, 0, cluster_num=kk)[0]
            pt.update(cpu_ms=dt * 1e3, cpu_frames_per_s=600 / dt)
            tot += dt
        cpu = dict(value=600 * len(SWEEP) / tot, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                   sample="the 9 sweep points once, full size, k-means++ sampled; 1 run", seconds_per_run=tot)
    _finish(ctx)
    if rank != 0:
        return
    ms_step = ms / args.steps
    frames = 600 * len(SWEEP)
    pk = _peaks()
    slow = max(pts, key=lambda p: p["gpu_ms"])
    line = dict(metric=METRIC, value=world * frames / (ms_step * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps, warmup=W_, ms_per_step=ms_step,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic", config=_config_dict(args), clocks=clocks,
                e2e=dict(value=world * frames / (ms_e2e / args.steps * 1e-3), unit=UNIT, h2d_bytes_per_step=sum(L.h2d_bytes() for L in lanes),
                         d2h_bytes_per_step=int(d2h[0]), ms_per_step=ms_e2e / args.steps),
                gpu_launches=int(launches), sweep=pts,
                roofline=dict(bound="hbm", kernel="whole path of the slowest point (latency-bound sizes)", achieved=None, peak=pk["hbm_gbs"], unit="GB/s",
                              frac=None, traffic=None, slowest_point=slow,
                              note="at 1k-16k tokens per stream every kernel is launch / latency bound; see the config2 line for the roofline"),
                pipeline=dict(cuda_graphs=use_graphs), all_gathered_stats_shape=list(gathered[0].shape))
    if cpu is not None:
        line["cpu_baseline"] = cpu
    _emit(line)


# ----------------------------------------------------------------------------------------------------------
# per-stage device times and the roofline (config-2 shape)
# ----------------------------------------------------------------------------------------------------------
def _stage_list(sc, d, ev_d, q_d):
    from cogstream_b200 import ops

    T, P, k = sc.T, sc.P, sc.k
    x2 = d.features.view(T, sc.K)

    def solve():
        ops.kmeans_solve(sc.G, d.timestamps, sc.km, n_init_given=0, exp_noise=None, reseed_pool=None, alpha=sc.cfg.alpha,
                         max_iteration=sc.cfg.max_iteration, tol=sc.cfg.tol, frames_per_warp=sc.cfg.solver_frames_per_warp,
                         rng_state=sc.rng_state)

    def select():
        ops.relevance_select(q_d.view(-1), ev_d, sc.G, sc.km.event_offsets, sc.km.event_order, sc.km.assignments, sc._ticket, k=k,
                             n_rep=sc.cfg.n_representatives, threshold=sc.cfg.relevance_threshold, bufs=sc._sel)

    def compact():
        ops.compact_stream(d.features, T, P, sc.mask, sc._sel["minor_flags"], min_tokens=sc.cfg.min_tokens, input_ids=d.input_ids,
                           image_token_index=sc.image_token_index, attention_mask=d.attention_mask, out=sc.out,
                           workspace=sc.compact_ws, bufs=sc._cmp)

    return [
        ("gram", lambda: ops.gram(x2, out=sc.G, workspace=sc.gram_ws)),
        ("kmeans_solve", solve),
        ("centroids", lambda: ops.centroid_gather(x2, sc.km.member_offsets, sc.km.members, sc.km.event_offsets, k,
                                                  sc.centroids, sc.packed)),
        ("relevance_select", select),
        ("pixel_diff", lambda: ops.pixel_diff(d.pixel_values, T, P, sc.W, out=sc.mask)),
        ("compact_stream", compact),
    ]


def _stage_times(sc, d, ev_d, q_d, reps=5):
    """Device time of every launcher of the path: each stage is launched `reps` times back to back between two CUDA events on
    the launching stream (so host launch latency is not in the number), after one in-order pass has set up its inputs."""
    T = sc.T
    stages = _stage_list(sc, d, ev_d, q_d)
    for _, fn in stages:  # in-order pass: every stage sees valid inputs
        fn()
    return {name: _ev_time(fn, reps) for name, fn in stages}


def _neighbour_ops(sc, d, reps=5):
    """The ops on either side of the timed path (NOT inside `value`): the fused
    embed_tokens + <image>-row replacement on the compacted prompt, the var-len mean pool of the k event-summary hidden
    states (k prompts of ~T*P/k image tokens, LLM width 3584), and the encoder tail (LayerNorm + bilinear merge) that produces
    the features.  HBM-bound; reported against the measured copy peak."""
    from cogstream_b200 import ops
    pk = _peaks()
    dev = d.features.device
    out = {}
    D = sc.D
    n_kept = int(sc.n_kept.item())
    vocab = 8192
    table = torch.randn(vocab, D, device=dev).to(torch.bfloat16)
    ids = torch.cat([torch.randint(0, vocab - 1, (14,), device=dev), torch.full((n_kept,), vocab - 1, device=dev, dtype=torch.int64),
                     torch.randint(0, vocab - 1, (23,), device=dev)])
    feats = sc.out[:n_kept]
    o = torch.empty(ids.numel(), D, dtype=torch.bfloat16, device=dev)
    ms = _ev_time(lambda: ops.embed_scatter(ids, vocab - 1, table, feats, out=o, check_count=False), reps)
    b = 2 * 2 * ids.numel() * D
    out["embed_scatter"] = dict(ms=round(ms, 5), rows=int(ids.numel()), D=D, algorithmic_bytes=b, gbs=round(b / ms / 1e6, 1),
                                frac=round(b / ms / 1e6 / pk["hbm_gbs"], 3))
    Dl, k = CFG2["Dl"], sc.k
    lens = [sc.T * sc.P // k + 40] * k
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    hidden = torch.randn(sum(lens), Dl, device=dev, dtype=torch.bfloat16)
    po = torch.empty(k, Dl, dtype=torch.bfloat16, device=dev)
    ms = _ev_time(lambda: ops.seq_mean_pool(hidden, cu, out=po), reps)
    b = 2 * sum(lens) * Dl
    out["seq_mean_pool"] = dict(ms=round(ms, 5), rows=int(sum(lens)), Dl=Dl, algorithmic_bytes=b, gbs=round(b / ms / 1e6, 1),
                                frac=round(b / ms / 1e6 / pk["hbm_gbs"], 3))
    del hidden, po
    # upstream: the encoder tail of one config-2 stream, post_layernorm + 2x2 bilinear merge in one pass
    m = CFG2["merge"]
    n_patch = sc.T * sc.P * m * m
    h = torch.randn(n_patch, D, device=dev).to(torch.bfloat16)
    g = torch.randn(D, device=dev).to(torch.bfloat16)
    bta = torch.randn(D, device=dev).to(torch.bfloat16)
    lo = torch.empty(n_patch // (m * m), D, dtype=torch.bfloat16, device=dev)
    ms = _ev_time(lambda: ops.ln_merge(h, m, g, bta, 1e-6, out=lo), reps)
    b = 2 * n_patch * D + 2 * (n_patch // (m * m)) * D
    out["ln_merge"] = dict(ms=round(ms, 5), rows=int(n_patch), D=D, merge=m, algorithmic_bytes=b, gbs=round(b / ms / 1e6, 1),
                           frac=round(b / ms / 1e6 / pk["hbm_gbs"], 3))
    return out


_KERNEL_OF_STAGE = {
    "gram": "gram_tc_kernel (+ gram_reduce_kernel)",
    "kmeans_solve": "kmeans_solver_kernel (k-means++ on CTA 0 + Lloyd loop, one cooperative launch)",
    "centroids": "centroid_gather_kernel",
    "pixel_diff": "pixel_diff_tma_kernel",
    "compact_stream": "compact_stream_coop_kernel + pool_emit_kernel",
}


def _roofline(stages, sc, ms_stream=None, ms_single=None):
    """`roofline` = the DOMINANT stage by measured device time, whatever it is (algorithmic bytes / flops as defined in
    DESIGN.md section 3.2), plus every wide kernel and the whole path."""
    pk = _peaks()
    T, P, D, k, K = sc.T, sc.P, sc.D, sc.k, sc.K
    n_kept = int(sc.n_kept.item())
    n_minor = int(sc.n_minor.item())
    BX = 2 * T * K
    alg_bytes = {
        "gram": BX,
        "kmeans_solve": 8 * T * T,  # the fp64 Gram matrix, read once (everything else is re-reads of it out of L2)
        "centroids": BX + 4 * k * K + (BX if sc.packed is not None else 0),
        "pixel_diff": 2 * T * P * sc.W,
        "compact_stream": 2 * 2 * n_kept * D + 2 * n_minor * P * D,
    }
    per = {n: dict(ms=stages[n], gbs=round(alg_bytes[n] / (stages[n] * 1e-3) / 1e9, 1),
                   frac=round(alg_bytes[n] / (stages[n] * 1e-3) / 1e9 / pk["hbm_gbs"], 4), algorithmic_bytes=alg_bytes[n])
           for n in alg_bytes if stages.get(n)}
    traffic = {}
    for name in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(tpath):
            traffic = json.load(open(tpath))
            break
    dom = max((n for n in stages if n in alg_bytes), key=lambda n: stages[n])
    gram_flops = T * (T + 1) * K  # 2 * (T(T+1)/2 unique entries) * K
    gram_tf = gram_flops / (stages["gram"] * 1e-3) / 1e12
    def _traffic(name):  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
        t = traffic.get(name)
        return t.get("total") if isinstance(t, dict) else t

    roof = dict(bound="hbm", kernel=_KERNEL_OF_STAGE.get(dom, dom), stage=dom, achieved=per[dom]["gbs"], peak=pk["hbm_gbs"], unit="GB/s",
                frac=per[dom]["frac"], traffic=_traffic(dom), algorithmic_bytes=alg_bytes[dom], ms=stages[dom])
    if dom == "kmeans_solve":
        roof["note"] = ("latency-bound: k-1 dependent k-means++ sampling rounds + the Lloyd iterations (grid barriers) on the L2-resident "
                        "2.9 MB Gram matrix; neither the HBM nor the tensor roofline is approached by construction.  Several streams in "
                        "flight hide it (`value`); it is what `single_stream` pays.")
    if dom == "gram":
        roof.update(bound="tensor", achieved=gram_tf, peak=pk["bf16_tflops"], unit="TFLOP/s", frac=gram_tf / pk["bf16_tflops"],
                    algorithmic_flops=gram_flops, hbm_view=per["gram"])
    roof["peaks"] = pk["source"]
    for n in per:
        per[n]["traffic"] = _traffic(n)
    roof["per_kernel_hbm"] = per
    roof["gram_tensor"] = dict(ms=stages["gram"], tflops=round(gram_tf, 1), frac_of_burst=round(gram_tf / pk["bf16_tflops"], 3),
                               algorithmic_flops=gram_flops)
    whole = sum(v for n, v in alg_bytes.items() if n != "kmeans_solve")
    total_ms = sum(stages.values())
    roof["whole_path"] = dict(algorithmic_bytes=whole, sum_stage_ms=round(total_ms, 4),
                              gbs=round(whole / (total_ms * 1e-3) / 1e9, 1), frac=round(whole / (total_ms * 1e-3) / 1e9 / pk["hbm_gbs"], 3))
    # the path as it is served: algorithmic bytes of one stream / steady-state device time per stream of the timed region
    # (`value`: several streams in flight) and / the latency of one stream alone.  north_star's ">= 60 % of HBM roofline".
    if ms_stream:
        g = whole / (ms_stream * 1e-3) / 1e9
        roof["whole_path"]["throughput"] = dict(ms_per_stream=round(ms_stream, 4), gbs=round(g, 1), frac=round(g / pk["hbm_gbs"], 3))
    if ms_single:
        g = whole / (ms_single * 1e-3) / 1e9
        roof["whole_path"]["single_stream"] = dict(ms=round(ms_single, 4), gbs=round(g, 1), frac=round(g / pk["hbm_gbs"], 3))
    return roof


_REAL_STDOUT = None


def _claim_stdout():
    """The driver reads ONE JSON line from stdout.  Native libraries (NCCL's version banner, ...) print to fd 1 directly, so the
    process-wide fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4", "config5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graphs", type=int, default=1, help="replay each phase as a CUDA graph (device-resident run)")
    ap.add_argument("--depth", type=int, default=6, help="independent streams kept in flight per GPU")
    ap.add_argument("--early-pixel", type=int, default=1, help="start the pixel-difference mask in phase 1, next to the k-means solver")
    ap.add_argument("--streams-per-step", type=int, default=32, help="config2: streams per rank per step")
    ap.add_argument("--e2e-steps", type=int, default=4, help="timed e2e steps (each moves streams_per_step x 759 MB over PCIe)")
    ap.add_argument("--segment", type=int, default=120, help="config4: new frames per question (120 = 60 s at 2 fps)")
    args = ap.parse_args()
    if os.environ.get("CSB_NOEARLY"):
        args.early_pixel = 0
    if args.impl == "reference":
        run_reference_arm(args)
        return
    ctx = Ctx(args)
    if args.workload in ("config2", "config3"):
        run_streams_workload(ctx, args)
    elif args.workload == "config4":
        run_config4(ctx, args)
    else:
        run_config5(ctx, args)


if __name__ == "__main__":
    main()
