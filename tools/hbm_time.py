"""CUDA-event timing of the HBM-bound kernels at LAP-3B B=32 shapes (L2 flushed between launches).

Prints algorithmic bytes / time for each kernel and the fraction of MEASURED_PEAKS.json's hbm_gbs.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200 import ops

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6540.8
try:
    peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dev = "cuda"
M, D, F = 22144, 2048, 16384
x = torch.randn(M, D, device=dev).bfloat16(); y = torch.empty_like(x); rstd = torch.rand(M, device=dev) + 0.5
scale = torch.randn(D, device=dev) * 0.1; dsc = torch.zeros(D, device=dev)
dy = torch.randn(M, D, device=dev).bfloat16(); dres = torch.randn(M, D, device=dev).bfloat16(); dx = torch.empty_like(x)
dact = torch.randn(M, F, device=dev).bfloat16(); gu = torch.randn(M, 2 * F, device=dev).bfloat16()
Ms, W = 16384, 1152
xs = torch.randn(Ms, W, device=dev).bfloat16(); ys = torch.empty_like(xs); mean = torch.zeros(Ms, device=dev); rs = torch.ones(Ms, device=dev)
dys = torch.randn(Ms, W, device=dev).bfloat16(); dxs = torch.empty_like(xs)
lsc, lbi = torch.ones(W, device=dev), torch.zeros(W, device=dev); dls, dlb = torch.zeros(W, device=dev), torch.zeros(W, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

cases = {
    "rmsnorm_fwd 22144x2048": (lambda: ops.rmsnorm_fwd(x, y, rstd, M, D, scale=scale), 2 * M * D * 2),
    "rmsnorm_bwd 22144x2048": (lambda: ops.rmsnorm_bwd(dy, x, scale, rstd, dres, dx, dsc, M, D), 4 * M * D * 2),
    "layernorm_fwd 16384x1152": (lambda: ops.layernorm_fwd(xs, lsc, lbi, ys, mean, rs, Ms, W), 2 * Ms * W * 2),
    "layernorm_bwd 16384x1152": (lambda: ops.layernorm_bwd(dys, xs, lsc, mean, rs, dres.view(-1)[: Ms * W].view(Ms, W), dxs, dls, dlb, Ms, W), 4 * Ms * W * 2),
    "geglu_bwd 22144x16384": (lambda: ops.geglu_bwd(dact, gu, M, F), 5 * M * F * 2),
}
out = {}
for name, (fn, nbytes) in cases.items():
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    gbs = nbytes / ms / 1e6
    out[name] = {"us": round(ms * 1e3, 1), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(gbs), "frac_of_peak": round(gbs / peak, 3)}
    print(f"{name:28s} {ms * 1e3:8.1f} us  {nbytes / 1e6:8.1f} MB  {gbs:6.0f} GB/s  {gbs / peak:.3f} of {peak}")
print(json.dumps(out))
