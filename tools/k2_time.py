"""K2 (fused SigLIP attention forward) stand-alone timing at the training / sweep shapes; LAPB_VIT_PAIR selects the kernel
(read once per process), so run it twice.  Also times the unfused GEMM + softmax + GEMM path on the same inputs."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200 import ops
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
out = {"pair_env": os.environ.get("LAPB_VIT_PAIR", "1")}
nh, hd = 16, 72
W = nh * hd
for Ni, Np in ((2, 256), (64, 256), (256, 256), (64, 729)):
    qkv = (torch.randn(Ni * Np, 3 * W, device="cuda") * 0.3).bfloat16()
    O = torch.empty(Ni * Np, W, device="cuda", dtype=torch.bfloat16)
    rec = {}
    if Np % 8 == 0:
        P = torch.empty(Ni, nh, Np, Np, device="cuda", dtype=torch.bfloat16)
        rec["fused_with_P_us"] = timeit(lambda: ops.vit_attn_fwd(qkv, O, P, Ni, nh, Np, hd, 0))
        qf = qkv.view(-1)
        def unfused():
            ops.gemm(qf, qf[W:], P, M=Np, N=Np, K=hd, lda=3 * W, ldb=3 * W, ldc=Np, batch_i=nh, batch_o=Ni,
                     a_bs=(hd, Np * 3 * W), b_bs=(hd, Np * 3 * W), c_bs=(Np * Np, nh * Np * Np))
            ops.vit_softmax_fwd(P, Ni * nh * Np, Np, Np, 0)
            ops.gemm(P, qf[2 * W:], O, M=Np, N=hd, K=Np, b_major=1, lda=Np, ldb=3 * W, ldc=W, batch_i=nh, batch_o=Ni,
                     a_bs=(Np * Np, nh * Np * Np), b_bs=(hd, Np * 3 * W), c_bs=(hd, Np * W))
        rec["unfused_us"] = timeit(unfused)
    rec["fused_no_P_us"] = timeit(lambda: ops.vit_attn_fwd(qkv, O, None, Ni, nh, Np, hd, 0))
    rec["flops"] = 4.0 * Ni * nh * Np * Np * hd
    rec["tflops_no_P"] = rec["flops"] / rec["fused_no_P_us"] / 1e6
    out[f"Ni{Ni}_Np{Np}"] = rec
print(json.dumps(out))
