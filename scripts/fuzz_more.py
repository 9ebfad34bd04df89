"""Randomised sweeps of three more code paths (test infrastructure, run by hand / under gpurun):
  gen    : the solver's built-in Philox generator at random shapes (fused launch, separate init kernel, the cooperative grid init of
           long streams; duplicated frames -> degenerate sampling rounds and re-seeds) against the CPU oracle fed with the SAME
           draws replayed on the host
  shard  : the frame-sharded Lloyd loop (ranks emulated in lock step) against the single-GPU solver, bit for bit
  cache  : StreamCache (incremental Gram + re-clustering after each appended segment) against the oracle on all frames so far
Usage: python scripts/fuzz_more.py [n_cases_per_sweep] [seed0]"""
import math, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import test_gpu_parity as tg  # _philox_host, _emulate_ranks
from cogstream_b200 import StreamCache, ops, synthetic
from oracle import cogstream_oracle as orc

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
fails = []
t_start = time.time()


def features(T, P, D, seed, dup):
    feats, _ = synthetic.make_features(T, P, D, seed)
    f3 = feats.view(T, P, D)
    if dup:
        g = torch.Generator().manual_seed(seed)
        src, dst = torch.randint(0, T, (T // 3,), generator=g), torch.randint(0, T, (T // 3,), generator=g)
        f3[dst] = f3[src]
    return f3


# ---- gen -----------------------------------------------------------------------------------------------------------
n_gen = 0
for case in range(n_cases):
    rnd = random.Random(seed0 + case)
    T = rnd.choice([40, 137, 300, 301, 602, 1023, 1024, 1025, 1400, 2100])
    k = rnd.choice([2, 5, 9, 20, 33, 64, 65, 130]); k = min(k, T // 2)
    P, D = rnd.choice([1, 2, 4]), rnd.choice([32, 64])
    dup = rnd.random() < 0.4
    heavy_dup = rnd.random() < 0.15  # only a handful of distinct frames: sampling rounds degenerate (sum == 0), many re-seeds
    ts_dtype = torch.float32 if rnd.random() < 0.25 else torch.float64
    f3 = features(T, P, D, seed0 + case, dup)
    if heavy_dup:
        f3[:] = f3[torch.randint(0, max(2, k // 2), (T,), generator=torch.Generator().manual_seed(case))]
    ts = (torch.arange(T, dtype=torch.float64) / rnd.choice([1.0, 2.0])).to(ts_dtype)
    G = ops.gram(f3.reshape(T, -1).to(dev))
    seed, off = rnd.getrandbits(62), rnd.getrandbits(20)
    state = torch.tensor([seed, off], dtype=torch.int64, device=dev)
    buf = ops.KMeansBuffers(T, k, ts_dtype, dev)
    if os.environ.get("FUZZ_VERBOSE"):
        print(f"gen case {case}: T={T} k={k} P={P} D={D} dup={dup} heavy_dup={heavy_dup} ts={ts_dtype}", flush=True)
    ops.kmeans_solve(G, ts.to(dev), buf, n_init_given=0, exp_noise=None, reseed_pool=None, rng_state=state)
    stt = ops.decode_status(buf.status)
    first = int((int(tg._philox_host(seed, off, 0xFFFFFFFF, [0])[0, 0]) * T) >> 32)
    q4 = (T + 3) // 4
    calls = tg._philox_host(seed, off, np.arange(k - 1)[:, None].repeat(q4, 1), np.arange(q4)[None, :].repeat(k - 1, 0))  # [k-1, q4, 4]
    u = (calls.reshape(k - 1, q4 * 4)[:, :T] >> np.uint32(8)).astype(np.float32)
    noise = torch.from_numpy(-np.log((u + np.float32(0.5)) * np.float32(2.0 ** -24)).astype(np.float32))
    reseeds = [int((int(v) * T) >> 32) for v in tg._philox_host(seed, off, 0xFFFFFFFE, np.arange(4096))[:, 0]]
    # the reference's random.randint calls, in order: the first centre (:41), one per degenerate sampling round (:57), then one per
    # empty cluster (:118) -- the solver takes the first from stream 0xFFFFFFFF and all the others, in that order, from 0xFFFFFFFE
    calls_made = {"n": 0}

    def replay_randint(a, b):
        calls_made["n"] += 1
        return first if calls_made["n"] == 1 else reseeds[calls_made["n"] - 2]

    class AfterInit:  # iterated once the init is over: continues where the degenerate rounds stopped
        def __iter__(self):
            return iter(reseeds[calls_made["n"] - 1:])

    info = orc.KMeansInfo()
    orig = random.randint
    try:
        random.randint = replay_randint
        _, ct_o, as_o = orc.kmeans_with_time_min_max(f3, ts, k, exp_noise=noise, reseed_indices=AfterInit(), info=info)
        ix = orc.KMeansInfo()
        calls_made["n"] = 0
        _, _, as_x = orc.kmeans_with_time_min_max(f3, ts, k, exp_noise=noise, reseed_indices=AfterInit(), info=ix, exact=True)
    finally:
        random.randint = orig
    if not torch.equal(as_x, as_o) or ix.init_indices != info.init_indices:
        continue  # decided by the reference's own fp32 rounding noise
    n_gen += 1
    good = (buf.init_indices.cpu().tolist() == info.init_indices and torch.equal(buf.assignments.cpu(), as_o)
            and stt["iterations"] == info.iterations and stt["degenerate_rounds"] == info.degenerate_rounds
            and stt["n_reseed_used"] == info.degenerate_rounds + len(info.reseed_indices)
            and state.tolist() == [seed, off + 1])
    if not good:
        ninit = sum(int(a != b) for a, b in zip(buf.init_indices.cpu().tolist(), info.init_indices))
        fails.append(("gen", case))
        print(f"gen case {case} FAILED: T={T} k={k} P={P} D={D} dup={dup} heavy={heavy_dup} {ts_dtype}: init diffs {ninit}, "
              f"assign diffs {int((buf.assignments.cpu() != as_o).sum())}, iters {stt['iterations']}/{info.iterations}, "
              f"pool draws {stt['n_reseed_used']}/{info.degenerate_rounds + len(info.reseed_indices)}, degenerate {stt['degenerate_rounds']}/{info.degenerate_rounds}")
print(f"gen: {n_gen} compared, {time.time() - t_start:.0f} s")

# ---- shard ---------------------------------------------------------------------------------------------------------
n_sh = 0
for case in range(n_cases):
    rnd = random.Random(seed0 + 1000 + case)
    T = rnd.choice([130, 257, 333, 600, 777, 1100, 2300])
    k = min(rnd.choice([3, 9, 23, 40, 64, 65, 129, 200]), T // 2)
    world = rnd.choice([2, 3, 4, 5, 8])
    ts_dtype = torch.float32 if rnd.random() < 0.25 else torch.float64
    dup = rnd.random() < 0.5
    f3 = features(T, rnd.choice([1, 2]), 32, seed0 + case, False)
    G = ops.gram(f3.reshape(T, -1).to(dev))
    ts = (torch.arange(T, dtype=torch.float64) / 2.0).to(ts_dtype).to(dev)
    g = torch.Generator().manual_seed(case)
    init = torch.randperm(T, generator=g)[:k]
    if dup:
        init[k // 2:] = init[0]
    pool = torch.randint(0, T, (k * 31 + 8,), generator=g).to(dev)
    use_rng = rnd.random() < 0.5
    ref = ops.KMeansBuffers(T, k, ts_dtype, dev)
    if use_rng:
        state = torch.tensor([rnd.getrandbits(60), rnd.getrandbits(16)], dtype=torch.int64, device=dev)
        ops.kmeans_solve(G, ts, ref, n_init_given=0, exp_noise=None, reseed_pool=None, rng_state=state.clone())
        ranks, n_coll = tg._emulate_ranks(T, k, ts, G, world, dev, rng_state=state)
    else:
        ref.init_indices.copy_(init)
        ops.kmeans_solve(G, ts, ref, n_init_given=k, exp_noise=None, reseed_pool=pool)
        ranks, n_coll = tg._emulate_ranks(T, k, ts, G, world, dev, init_indices=init, reseed_pool=pool)
    want = ops.decode_status(ref.status)
    n_sh += 1
    n = int(ref.member_offsets[-1])
    for km in ranks:
        got = ops.decode_status(km.buf.status)
        good = ((got["iterations"], got["converged"], got["n_reseed_used"], got["degenerate_rounds"], got["min_margin"]) ==
                (want["iterations"], want["converged"], want["n_reseed_used"], want["degenerate_rounds"], want["min_margin"])
                and torch.equal(km.buf.assignments, ref.assignments) and torch.equal(km.buf.centroid_times, ref.centroid_times)
                and torch.equal(km.buf.member_offsets, ref.member_offsets) and torch.equal(km.buf.members[:n], ref.members[:n])
                and torch.equal(km.buf.event_order, ref.event_order) and n_coll == 2 * want["iterations"])
        if not good:
            fails.append(("shard", case))
            print(f"shard case {case} FAILED on rank {km.rank}: T={T} k={k} world={world} {ts_dtype} dup={dup} rng={use_rng}: {got} vs {want}")
            break
print(f"shard: {n_sh} compared, {time.time() - t_start:.0f} s")

# ---- cache ---------------------------------------------------------------------------------------------------------
n_ca = 0
for case in range(max(3, n_cases // 3)):
    rnd = random.Random(seed0 + 2000 + case)
    P, D = rnd.choice([1, 3, 6]), rnd.choice([64, 256])
    segs = [rnd.choice([17, 60, 128, 129, 150, 200, 255]) for _ in range(rnd.randint(2, 4))]
    T_total = sum(segs)
    st = synthetic.make_stream(T_total, P, D, seed0 + case, C=8, fps=rnd.choice([1.0, 2.0]))
    f3 = st.features.view(T_total, P, D)
    cache = StreamCache(T_total + rnd.randint(0, 100), P, D, device=dev)
    g = torch.Generator().manual_seed(case)
    done = 0
    for seg in segs:
        cache.append(f3[done: done + seg].to(dev), st.timestamps[done: done + seg].to(dev))
        done += seg
        T = done
        k = math.ceil(T / 15)
        if T <= k or k < 2:
            continue
        init = torch.randperm(T, generator=g)[:k].tolist()
        pool = torch.randint(0, T, (256,), generator=g)
        buf, kk = cache.cluster(init_indices=init, reseed_pool=pool.to(dev))
        info, ix = orc.KMeansInfo(), orc.KMeansInfo()
        c, ct, a = orc.kmeans_with_time_min_max(f3[:T], st.timestamps[:T], k, init_indices=init, reseed_indices=pool.tolist(), info=info)
        _, _, ax = orc.kmeans_with_time_min_max(f3[:T], st.timestamps[:T], k, init_indices=init, reseed_indices=pool.tolist(), info=ix, exact=True)
        if not torch.equal(a, ax):
            continue
        n_ca += 1
        Gfull = ops.gram(f3[:T].reshape(T, -1).to(dev))
        good = (kk == k and torch.equal(buf.assignments.cpu(), a) and torch.equal(buf.centroid_times.cpu(), ct)
                and ops.decode_status(buf.status)["iterations"] == info.iterations
                and float((cache.gram - Gfull).abs().max() / Gfull.diagonal().max()) < 2e-6 and torch.equal(cache.gram, cache.gram.T)
                and tg.rel_err(cache.centroids(buf, k), c) < 1e-6)
        if not good:
            fails.append(("cache", case))
            print(f"cache case {case} FAILED at T={T} (segments {segs}, P={P}, D={D})")
            break
print(f"cache: {n_ca} compared, {time.time() - t_start:.0f} s")
print(f"fuzz_more: {len(fails)} failed {fails}")
sys.exit(1 if fails else 0)
