"""BASELINE.json configs[4]: SigLIP-So400m image-encoder-only forward throughput sweep, 224 and 384 px, batch 1..256.
(384 px = 27 x 27 = 729 patches with a random position table; K2 handles the three key chunks.)  One JSON record per
(resolution, images): ms, images/s, TFLOP/s (2*M*N*K of the tower) and the weight-streaming GB/s that bounds small batches."""
import dataclasses, gc, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200.config import get_config
from lap_b200.model import LAP, Staged

base = get_config("lap_libero").model
out = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for res in (224, 384):
    cfg = dataclasses.replace(base, image_size=res)
    model = LAP(cfg, seed=0)
    s, D, Np = cfg.siglip, cfg.gemma.width, cfg.num_patches
    W, F, L = s.width, s.mlp_dim, s.depth
    # per image: patch conv + L x (qkv, attention, out, fc1, fc2) + head
    flops = 2 * Np * (s.patch_size ** 2 * 3) * W + L * (2 * Np * W * 3 * W + 4 * Np * Np * W + 2 * Np * W * W + 4 * Np * W * F) \
        + 2 * Np * W * D
    wbytes = 2 * (L * (4 * W * W + 2 * W * F) + W * D)
    for n in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        if res == 384 and n > 128:
            continue  # the forward keeps every layer's activations (training layout): 256 images x 729 tokens does not fit
        B = max(1, n // 2)
        nimg = 2 * B
        imgs = [torch.rand(B, res, res, 3, device="cuda") * 2 - 1 for _ in range(2)]
        st = Staged(B=B, images=imgs, tokens=None, pm=None, par=None, pma=None, sm=None, sar=None)
        X0 = model.buf("sweep.X0", (B * cfg.prefix_len, D))
        for _ in range(2):
            model._siglip_fwd(st, X0, cfg.prefix_len, save=False)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            model._siglip_fwd(st, X0, cfg.prefix_len, save=False)
        reps = 20 if nimg <= 32 else 5
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rec = dict(resolution=res, patches=Np, images=nimg, ms=ms, images_per_s=nimg / ms * 1e3,
                   tflops=nimg * flops / ms / 1e9, weight_stream_gbs=wbytes / (ms * 1e-3) / 1e9,
                   gflops_per_image=flops / 1e9)
        out.append(rec); print(rec, flush=True)
        del g
        model._bufs.clear(); model._pool.clear()
        gc.collect(); torch.cuda.empty_cache()
    del model
    gc.collect(); torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/siglip_sweep.json", "w"), indent=1)
