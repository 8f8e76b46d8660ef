"""Small-M GEMMs of the batch-1 prefix pass: tile-shape / cta_group / split-K variants, CUDA-event timing (L2-warm, as in
the inference graph where the activations are L2-resident and the weights stream from HBM once)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200 import ops

shapes = [("down", 692, 2048, 16384), ("o", 692, 2048, 2048), ("qkv", 692, 2560, 2048), ("fc2", 512, 1152, 4304),
          ("out", 512, 1152, 1152), ("fc1", 512, 4304, 1152), ("vqkv", 512, 3456, 1152)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}
NW = 18  # rotate over 18 weight copies so that the weights come from HBM like in the real layer loop
for name, M, N, K in shapes:
    A = (torch.randn(M, K, device="cuda") * 0.1).bfloat16()
    Ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(NW)]
    C = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    C32 = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    R = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    variants = {"default": dict(), "cg1_bn128": dict(cta_group=1, block_n=128), "cg1_bn256": dict(cta_group=1, block_n=256),
                "cg2_bn128": dict(cta_group=2, block_n=128), "cg2_bn256": dict(cta_group=2, block_n=256)}
    out = {}
    for vn, kw in variants.items():
        try:
            for i in range(3):
                ops.gemm(A, Ws[i % NW], C, M=M, N=N, K=K, epi=ops.EPI_RESID, resid=R, **kw)
            torch.cuda.synchronize(); e0.record()
            for i in range(36):
                ops.gemm(A, Ws[i % NW], C, M=M, N=N, K=K, epi=ops.EPI_RESID, resid=R, **kw)
            e1.record(); torch.cuda.synchronize()
            out[vn] = e0.elapsed_time(e1) / 36 * 1e3
        except Exception as ex:
            out[vn] = repr(ex)[:80]
    for ks in (2, 3, 4, 6, 8):
        for cg, bn in ((1, 128), (2, 256)):
            vn = f"f32_splitk{ks}_cg{cg}_bn{bn}"
            try:
                for i in range(3):
                    ops.gemm(A, Ws[i % NW], C32, M=M, N=N, K=K, k_splits=ks, cta_group=cg, block_n=bn, accumulate=True)
                torch.cuda.synchronize(); e0.record()
                for i in range(36):
                    ops.gemm(A, Ws[i % NW], C32, M=M, N=N, K=K, k_splits=ks, cta_group=cg, block_n=bn, accumulate=True)
                e1.record(); torch.cuda.synchronize()
                out[vn] = e0.elapsed_time(e1) / 36 * 1e3
            except Exception as ex:
                out[vn] = repr(ex)[:80]
    res[name] = {"M": M, "N": N, "K": K, "us": out, "ideal_us_at_1400TF": 2 * M * N * K / 1.4e15 * 1e6}
    print(name, json.dumps(res[name]), flush=True)
json.dump(res, open("gpurun_out/gemm_small_m.json", "w"), indent=1)
