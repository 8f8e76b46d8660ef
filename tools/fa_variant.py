"""Time + checksum the fused Gemma attention forward of a given build: LAPB_LIB=<path to .so> python tools/fa_variant.py
(used to compare a variant build, e.g. -DFA_NPB=2, against the product library: identical checksums = identical bits)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lap_b200 import _lib
if os.environ.get("LAPB_LIB"):
    from pathlib import Path
    _lib.LIB_PATH = Path(os.environ["LAPB_LIB"]).resolve()
    _lib.needs_build = lambda: False
import torch
from lap_b200 import ops
out = {"lib": str(_lib.LIB_PATH.name)}
for tag, (B, T, Pn, S_len) in {"train": (32, 702, 692, 702), "ragged": (3, 333, 320, 301)}.items():
    NH, hd = 8, 256
    Tpad = (T + 63) // 64 * 64; W32 = Tpad // 32; R = T * NH
    g = torch.Generator(device="cuda").manual_seed(0)
    Q = (torch.randn(B, R, hd, device="cuda", generator=g) * 0.3).bfloat16()
    K = torch.randn(B, Tpad, hd, device="cuda", generator=g).bfloat16()
    V = torch.randn(B, Tpad, hd, device="cuda", generator=g).bfloat16()
    bits = torch.randint(-2**31, 2**31 - 1, (B, T, W32), dtype=torch.int32, device="cuda", generator=g) | 1
    P = torch.zeros(B, R, Tpad, dtype=torch.bfloat16, device="cuda")
    O0 = torch.zeros(B * Pn, NH * hd, dtype=torch.bfloat16, device="cuda")
    O1 = torch.zeros(B * (T - Pn), NH * hd, dtype=torch.bfloat16, device="cuda")
    run = lambda Pm: ops.fa_gemma_fwd(Q, K, V, bits, Pm, O0, O1, B, R, NH, T, S_len, Tpad, W32, Pn * NH, hd)
    run(P); torch.cuda.synchronize()
    cs = lambda t: [int(t.view(torch.int16).to(torch.int64).sum()), int((t.view(torch.int16).to(torch.int64) ** 2).sum() % (1 << 61))]
    out[tag] = {"O0": cs(O0), "O1": cs(O1), "P": cs(P), "finite": bool(torch.isfinite(O0.float()).all())}
    if tag == "train":
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for name, Pm in (("with_P_us", P), ("no_P_us", None)):
            ts = []
            for _ in range(12):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(Pm); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            ts.sort(); out[tag][name] = round(ts[len(ts) // 2] * 1e3, 1)
print(json.dumps(out))
