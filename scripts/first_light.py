"""GPU first-light: run each stage in its own subprocess with a timeout so a hang cannot eat the whole call.
Usage on the GPU box:  python scripts/first_light.py  (writes gpurun_out/first_light.log)"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

STAGES = {}


def stage(fn):
    STAGES[fn.__name__] = fn
    return fn


@stage
def s0_device():
    import torch
    from cogstream_b200 import ops
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), "cpus", os.cpu_count())
    ops.check_device()


@stage
def s1_gram_simt():
    import torch
    from cogstream_b200 import ops
    for T, K in [(64, 512), (150, 768), (600, 4096)]:
        x = torch.randn(T, K).bfloat16().cuda()
        G = ops.gram(x, impl=1)
        ref = x.double() @ x.double().T
        print("simt", T, K, float((G - ref).abs().max() / ref.diagonal().max()))


@stage
def s2_gram_tc():
    import torch
    from cogstream_b200 import ops
    for T, K in [(64, 64), (64, 512), (128, 128), (150, 768), (600, 4096), (300, 64), (129, 1600)]:
        x = torch.randn(T, K).bfloat16().cuda()
        G = ops.gram(x, impl=0)
        torch.cuda.synchronize()
        ref = x.double() @ x.double().T
        err = float((G - ref).abs().max() / ref.diagonal().max())
        print("tc", T, K, err, flush=True)
        if err > 1e-4:
            d = (G - ref).abs()
            bad = (d > 1e-4 * ref.diagonal().max()).nonzero()
            print("  bad entries:", bad.shape[0], bad[:8].tolist(), " G00", float(G[0, 0]), "ref00", float(ref[0, 0]))


@stage
def s3_kmeans_simt_gram():
    import torch
    from cogstream_b200 import kmeans_with_time_min_max
    gd = os.path.join(ROOT, "tests", "golden")
    for f in sorted(os.listdir(gd)):
        if not f.startswith("c"):
            continue
        g = torch.load(os.path.join(gd, f), weights_only=False)
        s = g["spec"]
        f3 = g["features"].view(s["T"], s["P"], s["D"]).cuda()
        cent, ct, a, info = kmeans_with_time_min_max(f3, g["timestamps"].cuda(), s["k"], init_indices=g["init"],
                                                     reseed_indices=[], gram_impl=1, return_info=True)
        same = torch.equal(a.cpu(), g["assignments"])
        print(f, "assign ok", same, "iters", info.iterations, g["iterations"], "margin", info.min_margin, g["min_margin"],
              "cent err", float((cent.cpu() - g["centroids"]).norm() / g["centroids"].norm()), flush=True)
        if not same:
            print("   ndiff", int((a.cpu() != g["assignments"]).sum()))


@stage
def s4_smoke():
    import __graft_entry__ as g
    g.smoke()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        STAGES[sys.argv[1]]()
        sys.exit(0)
    log = open(os.path.join(OUT, "first_light.log"), "w")
    for name in STAGES:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=240, cwd=ROOT)
            res = f"exit {p.returncode}"
            out = p.stdout[-6000:] + ("\nSTDERR:\n" + p.stderr[-3000:] if p.returncode else "")
        except subprocess.TimeoutExpired as e:
            res = "TIMEOUT"
            out = (e.stdout or b"").decode(errors="replace")[-4000:] if isinstance(e.stdout, bytes) else str(e.stdout)[-4000:]
        msg = f"===== {name}: {res} ({time.time() - t0:.1f}s)\n{out}\n"
        print(msg, flush=True)
        log.write(msg)
        log.flush()
