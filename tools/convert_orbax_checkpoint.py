"""Convert an OCDBT Orbax `params` item (the released LAP-3B / LAP-3B-Libero and openpi checkpoints) into a layout this
repo reads without orbax / tensorstore.  RUN THIS WHERE THE REFERENCE'S ENVIRONMENT IS INSTALLED (it imports orbax); the
build image has neither package, which is why the conversion is a separate one-off step.

  python tools/convert_orbax_checkpoint.py <ckpt_dir>/params <out_dir> [--format zarr|safetensors] [--dtype float32]

  zarr        : <out_dir>/params  in Orbax's own plain-directory layout (one zarr-v2 array per leaf, no OCDBT) — readable by
                lap_b200.orbax_io.read_params AND still by the reference's restore_params
  safetensors : <out_dir>/params.safetensors with '/'-joined reference key paths (lap_b200.checkpoint.load_tree)

The restore call is the reference's own `restore_params` (third_party/openpi/src/openpi/models/model.py:286-332) with
restore_type=np.ndarray, so the `value`-suffix handling is the reference's.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("params_dir")
    ap.add_argument("out_dir")
    ap.add_argument("--format", default="zarr", choices=["zarr", "safetensors"])
    ap.add_argument("--dtype", default=None, help="cast every leaf (e.g. float32); default: keep the stored dtype")
    args = ap.parse_args()
    try:
        from openpi.models import model as _model  # the reference's own reader
        params = _model.restore_params(args.params_dir, restore_type=np.ndarray)
    except ImportError:
        import orbax.checkpoint as ocp  # same calls as model.py:318-326 without the openpi package
        with ocp.PyTreeCheckpointer() as ckptr:
            meta = ckptr.metadata(args.params_dir)
            params = ckptr.restore(args.params_dir, ocp.args.PyTreeRestore(
                item={"params": meta["params"]},
                restore_args={"params": __import__("jax").tree.map(lambda _: ocp.RestoreArgs(restore_type=np.ndarray),
                                                                   meta["params"])}))["params"]
    from lap_b200 import orbax_io, params as P

    flat = P.from_nested(params) if not all(isinstance(k, str) and "/" in k for k in params) else params
    flat = {k[: -len("/value")] if k.endswith("/value") else k: np.asarray(v) for k, v in flat.items()}
    if args.dtype:
        flat = {k: v.astype(args.dtype) for k, v in flat.items()}
    os.makedirs(args.out_dir, exist_ok=True)
    if args.format == "zarr":
        orbax_io.write_params(os.path.join(args.out_dir, "params"), flat)
    else:
        import torch
        from lap_b200 import checkpoint as C
        C.save_tree(os.path.join(args.out_dir, "params.safetensors"), {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in flat.items()})
    print(f"wrote {len(flat)} tensors to {args.out_dir}")


if __name__ == "__main__":
    main()
