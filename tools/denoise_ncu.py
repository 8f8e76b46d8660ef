"""One fused sample_actions at LAP-3B expert size (for `ncu -k regex:denoise_loop -c 1`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200.config import LAPConfig
from lap_b200.data import synthetic_batch
from lap_b200.model import LAP
from lap_b200.observation import Observation

cfg = LAPConfig(paligemma_variant="mid_2b", action_expert_variant="gemma_300m", siglip_variant="tiny72/14", action_dim=7,
                action_horizon=10, max_token_len=180, enable_action_training=True, enable_image_augmentation=False,
                vocab_size=4096)
model = LAP(cfg, seed=0)
b = synthetic_batch(cfg, 1, step=0, with_langact=False)
model.use_cuda_graph = False
for _ in range(2):
    model.sample_actions(0, Observation.from_dict(b), num_steps=10, noise=b["noise"])
torch.cuda.synchronize()
