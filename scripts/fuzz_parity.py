"""Randomised parity sweep: N random small streams through the fused pipeline (C ABI) against the CPU oracle, bit-exact on the
assignments / iterations / minor list / mask / ids (cases that the reference's own fp32 rounding noise decides are skipped and
counted).  Test infrastructure: run by hand / under gpurun, not part of the driver's suites.
Usage: python scripts/fuzz_parity.py [n_cases] [seed0]"""
import math, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import StreamCompressor, StreamConfig, synthetic
from oracle import cogstream_oracle as orc

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda:0")
ok = skipped = tie_cases = 0
max_ct_ulps = 0.0
fails = []
t_start = time.time()
for case in range(n_cases):
    seed = seed0 + case
    rnd = random.Random(seed)
    T = rnd.choice([136, 137, 150, 199, 256, 257, 300, 333, 420, 511, 512, 513, 600, 640, 641, 700])
    P = rnd.choice([1, 2, 3, 5, 7, 8, 12, 16, 27])
    D = rnd.choice([64, 128, 1152])
    if T * P * D > 4e6:  # the CPU oracle runs twice per case
        P = max(1, int(4e6 // (T * D)))
    k = rnd.choice([None, None, None, 10, 11, 17, 32, 33, 64, 65, 100, 128, 129])
    iid = rnd.random() < 0.25
    dup = rnd.random() < 0.2       # duplicated frames: exact ties, empty clusters, re-seeds
    f32_time = rnd.random() < 0.2  # float32 timestamps
    fps = rnd.choice([1.0, 2.0, 3.0, 0.5])
    items = None
    if rnd.random() < 0.3:  # several visual items in one prompt: video segments, images, single-frame videos
        items, left = [], T
        while left > 0:
            kind = rnd.random()
            n = 1 if kind < 0.3 else min(left, rnd.randint(2, max(2, T // 2)))
            items.append((n, "image" if kind < 0.15 else "video"))
            left -= n
    st = synthetic.make_stream(T, P, D, seed, C=24, iid=iid, near_threshold_frac=0.1, fps=fps,
                               ts_dtype=torch.float32 if f32_time else torch.float64, items=items)
    if dup:
        f3 = st.features.view(T, P, D)
        src = torch.randint(0, T, (T // 5,), generator=torch.Generator().manual_seed(seed))
        dst = torch.randint(0, T, (T // 5,), generator=torch.Generator().manual_seed(seed + 1))
        f3[dst] = f3[src]
    kk = k if k is not None else math.ceil(T / 15)
    if kk >= T:
        continue
    ev, q = synthetic.make_embeddings(kk, 512, seed)
    g = torch.Generator().manual_seed(seed)
    init = torch.randperm(T, generator=g)[:kk]
    pool = torch.randint(0, T, (256,), generator=g)
    try:
        p1 = orc.phase1(st.features, T, st.timestamps, cluster_num=k, init_indices=init.tolist(), reseed_indices=pool.tolist())
        px = orc.phase1(st.features, T, st.timestamps, cluster_num=k, init_indices=init.tolist(), reseed_indices=pool.tolist(), exact=True)
    except Exception as exc:  # e.g. the re-seed pool ran out in the oracle
        skipped += 1
        print(f"case {case} skipped (oracle: {type(exc).__name__}: {exc})")
        continue
    if not torch.equal(px.assignments, p1.assignments) or px.info.iterations != p1.info.iterations:
        skipped += 1
        continue
    p2 = orc.phase2(st.features, p1, ev, q, st.pixel_values, st.grid_sizes, st.merge_sizes, st.modals, st.input_ids,
                    synthetic.IMAGE_TOKEN_INDEX, st.attention_mask)
    W = st.pixel_values.numel() // (T * P)
    sc = StreamCompressor(T, P, D, W, st.input_ids.numel(), device=dev, cfg=StreamConfig(cluster_num=k),
                          ts_dtype=torch.float32 if f32_time else torch.float64, items=items)
    d = st.to(dev)
    sc.phase1(d.features, d.timestamps, init_indices=init.to(dev), reseed_pool=pool.to(dev), pixel_values=d.pixel_values if case % 2 else None)
    sc.phase2(d.features, ev.to(dev), q.to(dev), d.pixel_values, d.input_ids, d.attention_mask)
    r = sc.results()
    ct_g, ct_o = r["centroid_times"].cpu(), p1.centroid_times
    ulp = torch.finfo(ct_o.dtype).eps * ct_o.abs().clamp(min=1e-30)
    ct_ulps = float(((ct_g.double() - ct_o.double()).abs() / ulp.double()).max())
    # representatives: torch.topk's order among exactly equidistant members (duplicated frames) is implementation-defined; the
    # library picks the lowest index.  Different indices with IDENTICAL features are the same choice; what follows from the
    # indices (minor list, mask, ids) is then not comparable and is skipped for that case.
    f3 = st.features.view(T, P, D)
    reps_same = all(torch.equal(a.cpu(), b) for a, b in zip(r["representatives"], p1.representatives))
    reps_equiv = reps_same
    if not reps_same:
        def canon(idx):
            rows = f3[idx.cpu().long()].reshape(len(idx), -1).float()
            order = sorted(range(len(idx)), key=lambda j: rows[j].tolist()[:16] + [float(rows[j].sum())])
            return rows[order]
        reps_equiv = all(len(a) == len(b) and torch.equal(canon(a), canon(b)) for a, b in zip(r["representatives"], p1.representatives))
    checks = dict(
        assignments=torch.equal(r["assignments"].cpu(), p1.assignments),
        iterations=r["status"]["iterations"] == p1.info.iterations,
        reseeds=r["status"]["n_reseed_used"] == len(p1.info.reseed_indices),
        centroid_times=ct_ulps <= 4.0,  # bit-exact for dyadic timestamps; the mean of 1/3-second or fp32 timestamps is summed in another order
        reps=reps_equiv,
        cos=torch.equal(r["cos"].cpu(), p2.cos),
    )
    if reps_same:
        checks.update(
            minor=r["minor"] == p2.minor,
            mask=torch.equal(r["mask"].cpu(), p2.mask),
            ids=torch.equal(r["input_ids"].cpu(), p2.input_ids),
            n_kept=r["n_kept"] == p2.features.shape[0],
        )
    else:
        tie_cases += 1
    max_ct_ulps = max(max_ct_ulps, ct_ulps)
    if all(checks.values()):
        ok += 1
    else:
        bad = [n for n, v in checks.items() if not v]
        fails.append((case, T, P, D, kk, iid, dup, f32_time, bad))
        print(f"case {case} FAILED: T={T} P={P} D={D} k={kk} iid={iid} dup={dup} f32_time={f32_time} items={items}: {bad} "
              f"(ct ulps {ct_ulps:.1f}, iters {r['status']['iterations']}/{p1.info.iterations}, margin gpu {r['status']['min_margin']:.2e} cpu {p1.info.min_margin:.2e})")
print(f"fuzz: {ok} ok ({tie_cases} of them with equivalent representatives among duplicated frames: downstream indices not compared), "
      f"{skipped} skipped (reference decided by its own rounding noise / oracle error), {len(fails)} failed; "
      f"centre times within {max_ct_ulps:.1f} ulp; {time.time() - t_start:.0f} s")
sys.exit(1 if fails else 0)
