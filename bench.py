#!/usr/bin/env python
"""Benchmark of the Visual Stream Compression hot path (BASELINE.json metric: frames/sec compressed).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the whole path (phase 1 + phase 2, the frozen-LLM forwards in between excluded, their
outputs being synthetic inputs) over one stream per rank of BASELINE config 2:
    T=600 frames (1 fps x 10 min), P=180 merged tokens/frame, D=1152 bf16 features, k=ceil(T/15)=40 events,
    pixel_values [T*720, 588] bf16, event/question embeddings [40,3584]/[1,3584] bf16, float64 timestamps.
Streams are independent, so N ranks run N streams with no data-path collective (weak scaling); the one
collective is the all-gather of the per-stream stats rows after the timed region's last step.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = through the public pipeline API
with pinned HOST inputs, H2D copies of every input and a D2H read of the result inside the timed region.
`--impl reference` times the CPU restatement of the reference (oracle/, kind "port": the reference is Python
and /root/reference is not on the GPU box) on the host cores, rank 0 only.
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
WORKLOAD = dict(T=600, P=180, D=1152, merge=2, C=588, Dl=3584, fps=1.0)


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
def _cpu_path_once(st, ev, q, k, seed):
    from cogstream_b200 import synthetic
    from oracle import cogstream_oracle as orc

    random.seed(seed)
    torch.manual_seed(seed)
    t0 = time.perf_counter()
    p1 = orc.phase1(st.features, st.T, st.timestamps, cluster_num=k)  # k-means++ sampled like the reference does
    p2 = orc.phase2(st.features, p1, ev, q, st.pixel_values, st.grid_sizes, st.merge_sizes, st.modals, st.input_ids,
                    synthetic.IMAGE_TOKEN_INDEX, st.attention_mask)
    dt = time.perf_counter() - t0
    return dt, p1, p2


def _make_cpu_sample(P_sample, seed=0):
    from cogstream_b200 import synthetic

    w = WORKLOAD
    st = synthetic.make_stream(w["T"], P_sample, w["D"], seed, fps=w["fps"], merge=w["merge"], C=w["C"])
    k = math.ceil(w["T"] / 15)
    ev, q = synthetic.make_embeddings(k, w["Dl"], seed)
    return st, ev, q, k


def cpu_baseline(budget_s: float, n_runs: int = 1):
    """Time the oracle on a bounded sample of config 2: same T, k, D; P reduced until one run fits the budget
    (the CPU cost is ~linear in P; the value reported is the one measured on the sample)."""
    torch.set_num_threads(os.cpu_count() or 1)
    w = WORKLOAD
    st, ev, q, k = _make_cpu_sample(6)
    t_small, _, _ = _cpu_path_once(st, ev, q, k, 0)
    t_small2, _, _ = _cpu_path_once(st, ev, q, k, 0)
    per_p = min(t_small, t_small2) / 6.0
    P_s = int(max(6, min(w["P"], budget_s / max(per_p, 1e-6) / max(n_runs, 1))))
    if P_s >= w["P"] * 0.9:
        P_s = w["P"]
    st, ev, q, k = _make_cpu_sample(P_s)
    times = []
    for r in range(n_runs):
        dt, p1, p2 = _cpu_path_once(st, ev, q, k, r)
        times.append(dt)
    t = statistics.median(times)
    sample = (f"1 stream of config 2 with T={w['T']} k={k} D={w['D']} and P={P_s} of {w['P']} tokens/frame "
              f"(CPU cost ~linear in P), k-means++ sampled, {n_runs} run(s), median")
    return dict(value=w["T"] / t, unit=UNIT, cores=torch.get_num_threads(), kind="port", sample=sample,
                seconds_per_run=t, value_full_P_estimate=w["T"] / (t * w["P"] / P_s)), times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total_budget = 150.0
    per_step = total_budget / max(args.steps + args.warmup, 1)
    torch.set_num_threads(os.cpu_count() or 1)
    w = WORKLOAD
    st, ev, q, k = _make_cpu_sample(6)
    _cpu_path_once(st, ev, q, k, 0)
    t_small, _, _ = _cpu_path_once(st, ev, q, k, 0)
    P_s = int(max(4, min(w["P"], per_step / max(t_small / 6.0, 1e-6))))
    if P_s >= w["P"] * 0.9:
        P_s = w["P"]
    st, ev, q, k = _make_cpu_sample(P_s)
    for i in range(args.warmup):
        _cpu_path_once(st, ev, q, k, 100 + i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        _cpu_path_once(st, ev, q, k, i)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    # `value` is for the FULL workload: when the bounded sample had to drop tokens per frame (large K), the measured step time is
    # scaled by P / P_s (the CPU path's cost is linear in P: cdist GEMMs over K = P*D, pixel mask and compaction over T*P rows)
    value = w["T"] / (dt * w["P"] / P_s)
    sample = (f"each step = 1 stream of config 2 with T={w['T']} k={k} D={w['D']} and P={P_s} of {w['P']} tokens/frame "
              f"(bounded so K+W steps end within minutes; value scaled by P/P_s to the full workload when P_s < P)")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=dt * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic", impl="reference",
                config=_config_dict(P_s),
                cpu_baseline=dict(value=value, unit=UNIT, cores=torch.get_num_threads(), kind="port", sample=sample,
                                  measured_sample_frames_per_s=w["T"] / dt),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    _emit(line)


def _config_dict(P=None):
    w = WORKLOAD
    return dict(workload="BASELINE config 2: single stream per GPU, 1 fps x 10 min", T=w["T"], P=P or w["P"], D=w["D"],
                k=math.ceil(w["T"] / 15), pixel_row=f"[T*{(P or w['P']) * w['merge'] ** 2}, {w['C']}] bf16", embed_dim=w["Dl"],
                timestamps="float64", init="k-means++ sampled on device (Exp(1) noise drawn in the timed region)",
                l2="inputs (757 MB/stream) exceed the 126 MB L2; no explicit flush", parallelism="independent streams, 1 per rank per step")


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from cogstream_b200 import STATS_FIELDS, StreamCompressor, _lib, gather_stats, synthetic

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; cogstream_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner / warnings never go to stdout
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    w = WORKLOAD
    T, P, D = w["T"], w["P"], w["D"]
    k = math.ceil(T / 15)

    # ---- synthetic stream of this rank (seed = rank), pinned on the host, resident copy on the device ----
    st = synthetic.make_stream(T, P, D, seed=rank, fps=w["fps"], merge=w["merge"], C=w["C"])
    ev, q = synthetic.make_embeddings(k, w["Dl"], seed=rank)
    host = st.pin()
    ev_h, q_h = ev.pin_memory(), q.pin_memory()
    d = st.to(dev)
    ev_d, q_d = ev.to(dev), q.to(dev)
    W = st.pixel_values.numel() // (T * P)
    torch.manual_seed(1234 + rank)
    random.seed(1234 + rank)
    h2d = sum(t.numel() * t.element_size() for t in (host.features, host.pixel_values, host.timestamps, host.input_ids,
                                                     host.attention_mask, ev_h, q_h))
    d2h = len(STATS_FIELDS) * 8 + T * P

    class Lane:
        """One in-flight step: its own CUDA stream, workspaces and (for e2e) staging / pinned result buffers.  Successive
        steps are independent streams of the workload, so `depth` of them are kept in flight (double buffering)."""

        def __init__(self):
            self.stream = torch.cuda.Stream(device=dev)
            self.sc = StreamCompressor(T, P, D, W, st.input_ids.numel(), device=dev)
            self.staged = False

        def stage(self):
            if self.staged:
                return
            self.e_feat = torch.empty_like(d.features); self.e_pix = torch.empty_like(d.pixel_values)
            self.e_ts = torch.empty_like(d.timestamps); self.e_ids = torch.empty_like(d.input_ids)
            self.e_am = torch.empty_like(d.attention_mask); self.e_ev = torch.empty_like(ev_d); self.e_q = torch.empty_like(q_d)
            self.r_stats = torch.empty(len(STATS_FIELDS), dtype=torch.int64).pin_memory()
            self.r_mask = torch.empty(T * P, dtype=torch.uint8).pin_memory()
            self.staged = True

    lanes = [Lane() for _ in range(max(1, args.depth))]
    sc = lanes[0].sc

    use_graphs = False
    launches_per_step = None
    if args.graphs:
        try:
            c0 = lib.csb_kernel_launches()
            for L in lanes:
                with torch.cuda.stream(L.stream):
                    L.sc.capture(d.features, d.timestamps, ev_d, q_d, d.pixel_values, d.input_ids, d.attention_mask)
            torch.cuda.synchronize()
            # capture() runs one eager pass + one captured pass per lane: launches of our kernels per step
            launches_per_step = (lib.csb_kernel_launches() - c0) // (2 * len(lanes))
            use_graphs = True
        except Exception as exc:  # capture is an optimisation of the launch path only
            print(f"bench.py: CUDA graph capture unavailable ({type(exc).__name__}: {exc}); running eager", file=sys.stderr)
            torch.cuda.synchronize()

    def step_resident(L):
        if use_graphs:
            L.sc.replay1()
            L.sc.replay2()
            return
        L.sc.phase1(d.features, d.timestamps)
        L.sc.phase2(d.features, ev_d, q_d, d.pixel_values, d.input_ids, d.attention_mask)

    def step_e2e(L):
        L.stage()
        L.e_feat.copy_(host.features, non_blocking=True)
        L.e_ts.copy_(host.timestamps, non_blocking=True)
        L.sc.phase1(L.e_feat, L.e_ts)
        L.e_pix.copy_(host.pixel_values, non_blocking=True)
        L.e_ids.copy_(host.input_ids, non_blocking=True)
        L.e_am.copy_(host.attention_mask, non_blocking=True)
        L.e_ev.copy_(ev_h, non_blocking=True)
        L.e_q.copy_(q_h, non_blocking=True)
        L.sc.phase2(L.e_feat, L.e_ev, L.e_q, L.e_pix, L.e_ids, L.e_am)
        L.r_stats.copy_(L.sc.stats_row(), non_blocking=True)
        L.r_mask.copy_(L.sc.mask, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, use_lanes):
        main = torch.cuda.current_stream(dev)
        for i in range(warmup * len(use_lanes)):
            L = use_lanes[i % len(use_lanes)]
            with torch.cuda.stream(L.stream):
                fn(L)
        barrier()
        l0 = lib.csb_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for L in use_lanes:
            L.stream.wait_event(e0)
        for i in range(steps):
            L = use_lanes[i % len(use_lanes)]
            with torch.cuda.stream(L.stream):
                fn(L)
        for L in use_lanes:
            done = torch.cuda.Event()
            done.record(L.stream)
            main.wait_event(done)
        e1.record(main)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.csb_kernel_launches() - l0
        if use_graphs and fn is step_resident:
            launches = launches_per_step * steps  # graph replays re-launch the captured kernels without the host-side counter
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total, launches = timed(step_resident, args.steps, max(args.warmup, 3), lanes)
    iters_row = sc.stats_row().cpu()
    ms_single, _ = timed(step_resident, args.steps, max(args.warmup, 3), lanes[:1])  # latency of one step, nothing overlapped
    ms_e2e, _ = timed(step_e2e, args.steps, max(args.warmup, 3), lanes)
    clocks = sampler.stop() if sampler else None  # sampled across the three timed regions (100 ms period)

    # the one collective: all-gather of the per-stream stats rows
    rows = gather_stats(sc.stats_row().unsqueeze(0))
    torch.cuda.synchronize()

    ms_step = ms_total / args.steps
    value = world * T / (ms_step * 1e-3)
    e2e_value = world * T / (ms_e2e / args.steps * 1e-3)

    # ---- per-stage device times (CUDA events between the ops of extra, untimed steps) + roofline ----
    stages = None
    roofline = None
    cpu = None
    if rank == 0:
        stages = _stage_times(sc, d, ev_d, q_d, reps=max(3, min(args.steps, 10)))
        roofline = _roofline(stages, sc, d, ms_step=ms_step, ms_single=ms_single / args.steps)
        roofline["neighbour_ops"] = _neighbour_ops(sc, d)
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = cpu_baseline(budget_s=20.0)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=_config_dict(), clocks=clocks,
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         ms_per_step=ms_e2e / args.steps),
                gpu_launches=int(launches), roofline=roofline, stages_ms=stages,
                pipeline=dict(depth=len(lanes), cuda_graphs=use_graphs, note="successive steps are independent streams; `depth` of them are in flight on "
                              "separate CUDA streams with their own workspaces", single_step_latency_ms=ms_single / args.steps,
                              frames_per_s_depth1=world * T / (ms_single / args.steps * 1e-3)),
                stream_stats=dict(zip(STATS_FIELDS, [int(v) for v in iters_row.tolist()])),
                all_gathered_stats_shape=list(rows.shape))
    if cpu is not None:
        line["cpu_baseline"] = cpu
    _emit(line)


def _stage_times(sc, d, ev_d, q_d, reps=5):
    """Device time of every launcher of the path: each stage is launched `reps` times back to back between two CUDA events on
    the launching stream (so host launch latency is not in the number), after one in-order pass has set up its inputs."""
    from cogstream_b200 import ops

    T, P, k = sc.T, sc.P, sc.k
    x2 = d.features.view(T, sc.K)
    noise = sc.noise[: k - 1]
    state = {}

    def solve():
        ops.kmeans_solve(sc.G, d.timestamps, sc.km, n_init_given=1, exp_noise=noise, reseed_pool=sc.pool, alpha=sc.cfg.alpha,
                         max_iteration=sc.cfg.max_iteration, tol=sc.cfg.tol)

    def reps_():
        state["rflags"] = ops.representatives_gram(sc.G, sc.km.event_offsets, sc.km.event_order, k, 2)[3]

    def relevance():
        state["mflags"] = ops.relevance_minor(q_d.view(-1), ev_d, sc.km.assignments, state["rflags"], sc._ticket, 0.45)[2]

    stages = [
        ("gram", lambda: ops.gram(x2, out=sc.G, workspace=sc.gram_ws)),
        ("rng_noise", lambda: noise.exponential_(1)),
        ("kmeans_solve", solve),
        ("centroid_gather", lambda: ops.centroid_gather(x2, sc.km.member_offsets, sc.km.members, sc.km.event_offsets, k,
                                                        sc.centroids, sc.packed)),
        ("representatives", reps_),
        ("relevance_minor", relevance),
        # the two launches of csb_pixel_mask timed separately: the HBM-bound difference kernel and the tiny per-frame fix-up
        ("pixel_diff", lambda: ops.pixel_diff(d.pixel_values, T, P, sc.W, out=sc.mask)),
        ("mask_finalize", lambda: ops.mask_finalize(sc.mask, T, P, minor_flags=state["mflags"])),
        ("compact_tokens", lambda: ops.compact_tokens(d.features, T, P, sc.mask, state["mflags"], out=sc.out, workspace=sc.compact_ws)),
        ("filter_ids", lambda: ops.filter_ids(d.input_ids, sc.image_token_index, sc.mask, d.attention_mask)),
    ]
    sc.km.init_indices[:1].fill_(random.randint(0, T - 1))
    sc.pool.random_(0, T)
    for _, fn in stages:  # in-order pass: every stage sees valid inputs
        fn()
    out = {}
    for name, fn in stages:
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = round(e0.elapsed_time(e1) / reps, 5)
    return out


def _neighbour_ops(sc, d, reps=5):
    """The two ops on either side of the timed path (NOT inside `value`): the fused
    embed_tokens + <image>-row replacement on the compacted prompt, and the var-len mean pool of the k event-summary hidden
    states (k prompts of ~T*P/k image tokens, LLM width 3584).  HBM-bound; reported against the measured copy peak."""
    from cogstream_b200 import ops
    pk = _peaks()
    dev = d.features.device
    out = {}

    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    D = sc.D
    n_kept = int(sc.n_kept.item())
    vocab = 8192
    table = torch.randn(vocab, D, device=dev).to(torch.bfloat16)
    ids = torch.cat([torch.randint(0, vocab - 1, (14,), device=dev), torch.full((n_kept,), vocab - 1, device=dev, dtype=torch.int64),
                     torch.randint(0, vocab - 1, (23,), device=dev)])
    feats = sc.out[:n_kept]
    o = torch.empty(ids.numel(), D, dtype=torch.bfloat16, device=dev)
    ms = timed(lambda: ops.embed_scatter(ids, vocab - 1, table, feats, out=o, check_count=False))
    b = 2 * 2 * ids.numel() * D
    out["embed_scatter"] = dict(ms=round(ms, 5), rows=int(ids.numel()), D=D, algorithmic_bytes=b, gbs=round(b / ms / 1e6, 1),
                                frac=round(b / ms / 1e6 / pk["hbm_gbs"], 3))
    Dl, k = WORKLOAD["Dl"], sc.k
    lens = [sc.T * sc.P // k + 40] * k
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    hidden = torch.randn(sum(lens), Dl, device=dev, dtype=torch.bfloat16)
    po = torch.empty(k, Dl, dtype=torch.bfloat16, device=dev)
    ms = timed(lambda: ops.seq_mean_pool(hidden, cu, out=po))
    b = 2 * sum(lens) * Dl
    out["seq_mean_pool"] = dict(ms=round(ms, 5), rows=int(sum(lens)), Dl=Dl, algorithmic_bytes=b, gbs=round(b / ms / 1e6, 1),
                                frac=round(b / ms / 1e6 / pk["hbm_gbs"], 3))
    return out


def _roofline(stages, sc, d, ms_step=None, ms_single=None):
    """Roofline of the dominant kernel (by measured time); algorithmic bytes / flops as defined in DESIGN.md."""
    pk = _peaks()
    T, P, D, k, K = sc.T, sc.P, sc.D, sc.k, sc.K
    n_kept = int(sc.n_kept.item())
    n_minor = int(sc.n_minor.item())
    BX = 2 * T * K
    alg_bytes = {
        "gram": BX,
        "centroid_gather": 2 * BX + 4 * k * K,
        "pixel_diff": 2 * T * P * sc.W,
        "compact_tokens": 2 * 2 * n_kept * D + 2 * n_minor * P * D,
    }
    hbm = {n: dict(ms=stages[n], gbs=alg_bytes[n] / (stages[n] * 1e-3) / 1e9, frac=alg_bytes[n] / (stages[n] * 1e-3) / 1e9 / pk["hbm_gbs"])
           for n in alg_bytes if stages.get(n)}
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath))
    # dominant kernel = the longest stage that has a hardware roofline (HBM or tensor); the k-means solver is latency-bound
    # by construction (L2-resident, 39 dependent sampling rounds) and is reported separately below
    dom = max((n for n in stages if n in alg_bytes), key=lambda n: stages[n])
    gram_flops = T * (T + 1) * K  # 2 * (T(T+1)/2 unique entries) * K
    gram_tf = gram_flops / (stages["gram"] * 1e-3) / 1e12
    if dom == "gram":
        roof = dict(bound="tensor", kernel="gram_tc_kernel", achieved=gram_tf, peak=pk["bf16_tflops"], unit="TFLOP/s",
                    frac=gram_tf / pk["bf16_tflops"], traffic=traffic.get("gram"),
                    hbm_view=dict(achieved=hbm["gram"]["gbs"], peak=pk["hbm_gbs"], unit="GB/s", frac=hbm["gram"]["frac"]))
    elif dom in hbm:
        kname = {"pixel_diff": "pixel_diff_tma_kernel", "centroid_gather": "centroid_gather_kernel",
                 "compact_tokens": "compact_scan_coop_kernel + gather_rows_kernel + pool_emit_kernel"}.get(dom, dom)
        roof = dict(bound="hbm", kernel=kname, stage=dom, achieved=hbm[dom]["gbs"], peak=pk["hbm_gbs"], unit="GB/s", frac=hbm[dom]["frac"],
                    traffic=traffic.get(dom))
    roof["peaks"] = pk["source"]
    roof["algorithmic_bytes"] = alg_bytes.get(dom)
    roof["latency_bound_stage"] = dict(kernel="kmeans_solver_kernel", ms=stages.get("kmeans_solve"), dram_traffic=traffic.get("kmeans_solve"),
                                       note="k-means++ + Lloyd on the L2-resident Gram matrix in one cooperative launch; no HBM / "
                                            "tensor roofline applies; hidden by keeping several streams in flight")
    roof["per_kernel_hbm"] = {n: dict(ms=v["ms"], gbs=round(v["gbs"], 1), frac=round(v["frac"], 3)) for n, v in hbm.items()}
    roof["gram_tensor"] = dict(ms=stages["gram"], tflops=round(gram_tf, 1), frac_of_burst=round(gram_tf / pk["bf16_tflops"], 3),
                               algorithmic_flops=gram_flops)
    whole = sum(alg_bytes.values())
    total_ms = sum(stages.values())
    roof["whole_path"] = dict(algorithmic_bytes=whole, sum_stage_ms=round(total_ms, 4),
                              gbs=round(whole / (total_ms * 1e-3) / 1e9, 1), frac=round(whole / (total_ms * 1e-3) / 1e9 / pk["hbm_gbs"], 3))
    # the path as it is served: algorithmic bytes of one stream / steady-state device time per stream of the timed region
    # (`value`: several streams in flight) and / the latency of one stream alone.  north_star's ">= 60 % of HBM roofline".
    if ms_step:
        g = whole / (ms_step * 1e-3) / 1e9
        roof["whole_path"]["throughput"] = dict(ms_per_stream=round(ms_step, 4), gbs=round(g, 1), frac=round(g / pk["hbm_gbs"], 3))
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graphs", type=int, default=1, help="replay each phase as a CUDA graph (device-resident run)")
    ap.add_argument("--depth", type=int, default=6, help="steps (independent streams) kept in flight per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
