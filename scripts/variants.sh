mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err; python - <<PY
import json
j=json.loads(open('gpurun_out/r2g_$name.json').read().strip().splitlines()[-1])
print('$name', 'value %.0f'%j['value'], 'ms/stream %.4f'%j['pipeline']['ms_per_stream'], 'single %.4f'%j['single_stream']['ms'], 'solve %.4f'%j['stages_ms']['kmeans_solve'])
PY
}
run base A=1
run nodraws CSB_PAR_DRAWS=0
run nodraws_fpw2 CSB_PAR_DRAWS=0 CSB_SOLVER_FPW=2
run nodraws_noearly CSB_PAR_DRAWS=0 CSB_NOEARLY=1
timeout 600 python -m pytest tests -m gpu -q -k "encoder_tail or multi_item or relevance_select or kmeanspp_approx" > gpurun_out/r2g_tests.log 2>&1; tail -4 gpurun_out/r2g_tests.log
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from cogstream_b200 import ops
dev = torch.device('cuda:0')
T, G, C = 600, 720, 1152
h = torch.randn(T * G, C, device=dev).to(torch.bfloat16)
w = torch.ones(C, device=dev, dtype=torch.bfloat16); b = torch.zeros(C, device=dev, dtype=torch.bfloat16)
out = torch.empty(T * G // 4, C, device=dev, dtype=torch.bfloat16)
for _ in range(3): ops.ln_merge(h, 2, w, b, 1e-6, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.ln_merge(h, 2, w, b, 1e-6, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = h.numel() * 2 + out.numel() * 2
print(f"ln_merge config-2 size: {ms*1e3:.1f} us, {byt/1e6:.0f} MB algorithmic -> {byt/ms/1e6:.0f} GB/s ({byt/ms/1e6/6545.6:.3f} of the measured HBM peak)")
# the reference's op chain on the same device
import importlib
sys.path.insert(0, '.')
from oracle import cogstream_oracle as orc
grid = torch.tensor([[T, 24, 30]], device=dev); merge = torch.tensor([2], device=dev)
for _ in range(2): orc.encoder_tail(h, grid, merge, w, b, 1e-6)
torch.cuda.synchronize(); e0.record()
for _ in range(5): orc.encoder_tail(h, grid, merge, w, b, 1e-6)
e1.record(); torch.cuda.synchronize()
print(f"reference op chain (torch CUDA ops): {e0.elapsed_time(e1)/5*1e3:.1f} us")
PY
