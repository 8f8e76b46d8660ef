"""Bottleneck attribution of the fused Gemma attention forward (K1) at the LAP-3B training shape: times the kernel of a
-DLAPB_FA_KNOBS build (lap_b200/csrc/liblapb200_knobs.so, see the FA_KNOB comment in fa_gemma.cu) with one pipeline stage
switched off at a time.  Outputs of knob runs are wrong by construction; only the durations mean something.

build:  cd lap_b200/csrc && nvcc <NVCC_FLAGS of _lib.py> -I ../../include -DLAPB_FA_KNOBS -c fa_gemma.cu -o fa_gemma_knobs.o \
        && nvcc -shared -o liblapb200_knobs.so api.o gemm.o elementwise.o attention.o loss.o optimizer.o skinny.o \
           fa_gemma_knobs.o denoise.o -gencode arch=compute_100a,code=sm_100a
run:    python tools/fa_knobs.py > gpurun_out/fa_knobs.json
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lap_b200 import _lib
_lib.LIB_PATH = _lib.CSRC / "liblapb200_knobs.so"
_lib.needs_build = lambda: False
import torch
from lap_b200 import ops

B, T, NH, hd, Pn = 32, 702, 8, 256, 692
Tpad = 704; W32 = Tpad // 32; R = T * NH
g = torch.Generator(device="cuda").manual_seed(0)
Q = (torch.randn(B, R, hd, device="cuda", generator=g) * 0.1).bfloat16()
K = torch.randn(B, Tpad, hd, device="cuda", generator=g).bfloat16()
V = torch.randn(B, Tpad, hd, device="cuda", generator=g).bfloat16()
bits = torch.full((B, T, W32), -1, dtype=torch.int32, device="cuda")
P = torch.empty(B, R, Tpad, dtype=torch.bfloat16, device="cuda")
O0 = torch.empty(B * Pn, NH * hd, dtype=torch.bfloat16, device="cuda")
O1 = torch.empty(B * (T - Pn), NH * hd, dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
NAMES = {1: "exp2->fmul", 2: "no pass 1", 4: "no PV MMAs", 8: "no P write/store", 16: "no pass-2 S MMAs", 32: "no mask words"}
CASES = [0, 1, 32, 2, 4, 8, 16, 4 | 8, 4 | 16, 1 | 32, 2 | 4 | 16, 1 | 2 | 32, 1 | 4 | 8 | 16 | 32, 1 | 2 | 4 | 8 | 16 | 32]


def run(Pm):
    ops.fa_gemma_fwd(Q, K, V, bits, Pm, O0, O1, B, R, NH, T, T, Tpad, W32, Pn * NH, hd)


out = []
for knobs in CASES:
    os.environ["LAPB_FA_KNOBS"] = str(knobs)
    row = {"knobs": knobs, "off": [n for k, n in NAMES.items() if knobs & k]}
    for name, Pm in (("with_P_us", P), ("no_P_us", None)):
        for _ in range(2):
            run(Pm)
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(Pm); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ts.sort()
        row[name] = round(ts[len(ts) // 2] * 1e3, 1)
    out.append(row)
    print(json.dumps(row), file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))
