"""BASELINE.json configs[4]: SigLIP-So400m image-encoder-only forward throughput sweep (224 px, batch 1..256)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lap_b200.config import get_config
from lap_b200.model import LAP, Staged
tc = get_config("lap_libero"); cfg = tc.model
model = LAP(cfg, seed=0)
D = cfg.gemma.width; Np = cfg.num_patches
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = []
flops_per_image = 220.2e9
for n in (1, 2, 4, 8, 16, 32, 64, 128, 256):
    # n images = n/2 "samples" x 2 cameras (or 1 sample x 1 camera for n = 1 -> use 2 cams, count 2 images)
    B = max(1, n // 2); nimg = 2 * B
    imgs = [torch.rand(B, 224, 224, 3, device="cuda") * 2 - 1 for _ in range(2)]
    st = Staged(B=B, images=imgs, tokens=None, pm=None, par=None, pma=None, sm=None, sar=None)
    X0 = model.buf("sweep.X0", (B * cfg.prefix_len, D))
    for _ in range(2): model._siglip_fwd(st, X0, cfg.prefix_len)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        model._siglip_fwd(st, X0, cfg.prefix_len)
    torch.cuda.synchronize(); e0.record()
    reps = 20 if nimg <= 32 else 5
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rec = dict(images=nimg, ms=ms, images_per_s=nimg / ms * 1e3, tflops=nimg * flops_per_image / ms / 1e9,
               weight_stream_gbs=0.83e9 / (ms * 1e-3) / 1e9)
    out.append(rec); print(rec, flush=True)
    model._bufs = {k: v for k, v in model._bufs.items() if not k.startswith("img.") and not k.startswith("sweep.")}
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/siglip_sweep.json", "w"), indent=1)
