import numpy as np
import torch


def obs_for_oracle(b, langact=True):
    t = lambda x: torch.from_numpy(np.asarray(x))
    return dict(images={k: t(v) for k, v in b["image"].items()}, image_masks={k: t(v) for k, v in b["image_mask"].items()},
                tokenized_prompt=t(b["tokenized_prompt"]), tokenized_prompt_mask=t(b["tokenized_prompt_mask"]),
                tokenized_langact_mask=t(b["tokenized_langact_mask"]) if (langact and "tokenized_langact_mask" in b) else None,
                token_loss_mask=t(b["token_loss_mask"]), sample_mask=t(b["sample_mask"]))


def rel_err(a, b):
    a = torch.as_tensor(a).detach().float().cpu()
    b = torch.as_tensor(b).detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
