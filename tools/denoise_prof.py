"""Per-phase profile of the persistent denoise-loop kernel (K10) at LAP-3B expert size, plus event timing of the
fused loop vs the kernel-per-op loop on the same prefix cache."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200.config import LAPConfig, get_config
from lap_b200.data import synthetic_batch
from lap_b200.model import LAP
from lap_b200.observation import Observation

full = len(sys.argv) > 1 and sys.argv[1] == "full"
cfg = get_config("lap_libero").model if full else LAPConfig(
    paligemma_variant="mid_2b", action_expert_variant="gemma_300m", siglip_variant="tiny72/14", action_dim=7,
    action_horizon=10, max_token_len=180, enable_action_training=True, enable_image_augmentation=False, vocab_size=4096)
model = LAP(cfg, seed=0)
b = synthetic_batch(cfg, 1, step=0, with_langact=False)
obs = Observation.from_dict(b)
model.use_cuda_graph = False
out = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for fused in (False, True):
    model.use_denoise_megakernel = fused
    for steps in (1, 10):
        for _ in range(2):
            model.sample_actions(0, obs, num_steps=steps, noise=b["noise"])
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            model.sample_actions(0, obs, num_steps=steps, noise=b["noise"])
        e1.record(); torch.cuda.synchronize()
        out[f"{'fused' if fused else 'per_op'}_steps{steps}_ms"] = e0.elapsed_time(e1) / 5
model.denoise_profile = True
model.sample_actions(0, obs, num_steps=10, noise=b["noise"])
torch.cuda.synchronize()
prof = model._bufs["dn.prof"].cpu().tolist()
names = ["prologue", "action_in", "P1 work", "P1 barrier", "P2 work", "P2 barrier", "P2b work", "P2b barrier", "P3 work",
         "P3 barrier", "P4 work", "P4 barrier", "P5 work", "P5 barrier", "final"]
out["phase_us_total_10_steps"] = {n: v / 1e3 for n, v in zip(names, prof)}
out["phase_us_per_layer_step"] = {n: v / 1e3 / (10 * cfg.gemma.depth) for n, v in zip(names[2:14], prof[2:14])}
sub = {16: "P1 stage+wait", 17: "P1 norm", 18: "P1 mma", 19: "P1 epilogue", 20: "P1 kv preload", 2: "P1 prefetch",
       21: "P2b loop", 22: "P2b load_w", 6: "P2b prefetch", 24: "P2 stage+wait", 25: "P2 rope", 26: "P2 S mma", 27: "P2 softmax",
       4: "P2 PV+store", 28: "P4 stage+wait", 29: "P4 norm", 30: "P4 passes", 10: "P4 w5+prefetch"}
out["sub_us_per_layer_step"] = {n: prof[i] / 1e3 / (10 * cfg.gemma.depth) for i, n in sub.items()}
out["error_flag"] = model.denoise_error_flag()
print(json.dumps(out, indent=1))
