"""Per-phase profile of the persistent denoise-loop kernel (K10) at LAP-3B expert size (`full`) or a mid-size model,
plus event timing of the fused loop (profiling off) on the same prefix cache.

Every CTA's thread 0 accumulates globaltimer nanoseconds per slot in shared memory (denoise.cu `tick`), so the table
shows, per slot, CTA 0 and the min / median / max over the grid: a slot whose max is far above its median has
stragglers; the barrier slots (`wait`) are smallest on the CTA that arrives last.

usage: python tools/denoise_prof.py [full] [--no-per-op]     (LAPB_DENOISE_FLAGS / LAPB_DENOISE_CTAS select the variant)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200.config import LAPConfig, get_config
from lap_b200.data import synthetic_batch
from lap_b200.model import LAP
from lap_b200.observation import Observation

full = "full" in sys.argv[1:]
per_op = "--no-per-op" not in sys.argv[1:]
cfg = get_config("lap_libero").model if full else LAPConfig(
    paligemma_variant="mid_2b", action_expert_variant="gemma_300m", siglip_variant="tiny72/14", action_dim=7,
    action_horizon=10, max_token_len=180, enable_action_training=True, enable_image_augmentation=False, vocab_size=4096)
model = LAP(cfg, seed=0)
b = synthetic_batch(cfg, 1, step=0, with_langact=False)
obs = Observation.from_dict(b)
model.use_cuda_graph = False
out = {"flags": os.environ.get("LAPB_DENOISE_FLAGS", "0"), "ctas": os.environ.get("LAPB_DENOISE_CTAS")}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for fused in ((False, True) if per_op else (True,)):
    model.use_denoise_megakernel = fused
    for steps in (1, 10):
        for _ in range(2):
            model.sample_actions(0, obs, num_steps=steps, noise=b["noise"])
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            model.sample_actions(0, obs, num_steps=steps, noise=b["noise"])
        e1.record(); torch.cuda.synchronize()
        out[f"{'fused' if fused else 'per_op'}_steps{steps}_ms"] = e0.elapsed_time(e1) / 5
model.use_denoise_megakernel = True
model.denoise_profile = True
model.sample_actions(0, obs, num_steps=10, noise=b["noise"])
torch.cuda.synchronize()
prof = model._bufs["dn.prof"].cpu().view(-1, 32).double()
used = prof.sum(1) > 0
prof = prof[used]
out["ctas_profiled"] = int(used.sum())
LS = 10 * cfg.gemma.depth
# slot -> name, in program order; the interval a slot measures ends at the tick of that number (denoise.cu)
slots = [(15, "prologue T1 time_mlp_in"), (23, "prologue T2 time_mlp_out"), (0, "prologue T3 modulation + init"),
         (1, "action_in (per step)"),
         (16, "P1 stage XE+mod"), (17, "P1 norm"), (18, "P1 mma (all passes)"), (19, "P1 epilogues"), (20, "P1 K/V preload issue"),
         (2, "P1 -"), (3, "P1->P2 wait"),
         (24, "P2 stage q"), (25, "P2 rope"), (26, "P2 S mma"), (27, "P2 softmax"), (4, "P2 PV+store"), (5, "P2->P2b wait"),
         (21, "P2b combine"), (22, "P2b w3 issue"), (6, "P2b -"), (7, "P2b->P3 wait"),
         (8, "P3 o-proj"), (9, "P3->P4 wait"),
         (28, "P4 stage XE1"), (29, "P4 norm"), (30, "P4 passes"), (10, "P4 w5 issue"), (11, "P4->P5 wait"),
         (12, "P5 down + w1 issue"), (13, "P5->P1 wait"), (14, "final (per step)")]
tab = {}
tot0 = 0.0
for i, name in slots:
    col = prof[:, i] / 1e3
    div = 1 if i in (0, 15, 23) else (10 if i in (1, 14) else LS)
    tab[name] = {"cta0": float(col[0]) / div, "min": float(col.min()) / div, "median": float(col.median()) / div,
                 "max": float(col.max()) / div}
    if i not in (0, 1, 14, 15, 23):
        tot0 += float(col[0]) / div
out["us_per_layer_step"] = tab
out["cta0_layer_step_us"] = tot0
out["error_flag"] = model.denoise_error_flag()
print(json.dumps(out, indent=1))
