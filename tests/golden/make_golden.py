"""Generates tests/golden/*.npz from the ORACLE (oracle/lap_oracle.py).

The reference (JAX) cannot run in this environment, so these fixtures pin the oracle itself (against regressions) and
give the GPU tests fixed inputs/outputs; they are NOT reference outputs.  Re-run: python tests/golden/make_golden.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from lap_b200 import params as P
from lap_b200.config import get_config
from lap_b200.data import synthetic_batch
from oracle import lap_oracle as O

OUT = os.path.dirname(os.path.abspath(__file__))


def obs_of(b, langact=True):
    t = lambda x: torch.from_numpy(np.asarray(x))
    return dict(images={k: t(v) for k, v in b["image"].items()}, image_masks={k: t(v) for k, v in b["image_mask"].items()},
                tokenized_prompt=t(b["tokenized_prompt"]), tokenized_prompt_mask=t(b["tokenized_prompt_mask"]),
                tokenized_langact_mask=t(b["tokenized_langact_mask"]) if langact else None,
                token_loss_mask=t(b["token_loss_mask"]), sample_mask=t(b["sample_mask"]))


def main():
    for name, B in (("debug_tiny", 3), ("debug_small", 2)):
        tc = get_config(name); cfg = tc.model
        ref = P.init_reference_params(cfg, 7, reference_zero_init=False)
        b = synthetic_batch(cfg, B, step=11)
        t = lambda x: torch.from_numpy(np.asarray(x))
        out = {}
        for bf in (True, False):
            loss, m, aux = O.compute_loss(ref, cfg, obs_of(b), t(b["actions"]), t(b["noise"]), t(b["time"]), bf16=bf, return_aux=True)
            tag = "bf16" if bf else "f32"
            out[f"loss_{tag}"] = loss.numpy()
            for k, v in m.items(): out[f"{k}_{tag}"] = v.numpy()
            out[f"v_t_{tag}"] = aux["v_t"].numpy()
            out[f"actions_{tag}"] = O.sample_actions(ref, cfg, obs_of(b, langact=False), t(b["noise"]), num_steps=10, bf16=bf).numpy()
            if bf:
                out["mask"] = np.packbits(aux["mask"].numpy(), axis=-1)
                out["positions"] = aux["positions"].numpy()
        np.savez_compressed(os.path.join(OUT, f"{name}_B{B}.npz"), **out)
        print(name, {k: (v.shape, float(np.asarray(v, dtype=np.float64).ravel()[0])) for k, v in out.items() if k.startswith("loss")})


if __name__ == "__main__":
    main()
