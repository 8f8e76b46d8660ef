"""Event timing of the fused Gemma attention forward (K1) at the LAP-3B training shape (B=32, T=702, 8 heads x 256)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200 import ops
B, T, NH, hd, Pn = 32, 702, 8, 256, 692
Tpad = 704; W32 = Tpad // 32; R = T * NH
g = torch.Generator(device="cuda").manual_seed(0)
Q = (torch.randn(B, R, hd, device="cuda", generator=g) * 0.1).bfloat16()
K = (torch.randn(B, Tpad, hd, device="cuda", generator=g)).bfloat16()
V = (torch.randn(B, Tpad, hd, device="cuda", generator=g)).bfloat16()
bits = torch.full((B, T, W32), -1, dtype=torch.int32, device="cuda")
P = torch.empty(B, R, Tpad, dtype=torch.bfloat16, device="cuda")
O0 = torch.empty(B * Pn, NH * hd, dtype=torch.bfloat16, device="cuda")
O1 = torch.empty(B * (T - Pn), NH * hd, dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for name, Pm in (("with_P", P), ("no_P", None)):
    for _ in range(3):
        ops.fa_gemma_fwd(Q, K, V, bits, Pm, O0, O1, B, R, NH, T, T, Tpad, W32, Pn * NH, hd)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.fa_gemma_fwd(Q, K, V, bits, Pm, O0, O1, B, R, NH, T, T, Tpad, W32, Pn * NH, hd); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    us = ts[len(ts) // 2] * 1e3
    flops = 4.0 * B * R * Tpad * hd  # QK^T + PV (the recomputed QK^T of pass 1 is not counted)
    out[name] = {"us": us, "algorithmic_TFLOPs": flops / us / 1e6, "executed_TFLOPs": 1.5 * flops / us / 1e6}
print(json.dumps(out, indent=1))
