"""Time csb_ln_merge at the config-2 size (432,000 patches x 1152 channels, merge 2) and print a checksum of the output.
CSB_LN_COLUMNWISE=1 selects the column-wise kernel (validation)."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from cogstream_b200 import ops
dev = torch.device("cuda:0")
T, G, C = 600, 720, 1152
g = torch.Generator(device=dev).manual_seed(0)
h = (torch.randn(T * G, C, device=dev, generator=g) * 3 + 0.7).to(torch.bfloat16)
w = torch.randn(C, device=dev, generator=g).to(torch.bfloat16)
b = torch.randn(C, device=dev, generator=g).to(torch.bfloat16)
for m in (2, 1, 3):
    n = T * G // (m * m) * (m * m)
    out = torch.empty(n // (m * m), C, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.ln_merge(h[:n], m, w, b, 1e-6, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.ln_merge(h[:n], m, w, b, 1e-6, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    rows_read = n if m % 2 == 0 else n // (m * m)  # odd merge sizes read the central patch only
    rows_read = n // (m * m) * (4 if m % 2 == 0 else 1)
    byt = rows_read * C * 2 + out.numel() * 2
    digest = hashlib.sha1(out.view(torch.int16).cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"merge {m}: {ms*1e3:7.1f} us, {byt/1e6:6.0f} MB -> {byt/ms/1e6:5.0f} GB/s ({byt/ms/1e6/6545.6:.3f} of the measured HBM peak)  sha1 {digest}")
