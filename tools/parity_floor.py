"""How far apart are two CORRECT bf16 implementations of the full-size LAP-3B forward?  (CPU only, ~3 minutes.)

The bf16-emulating oracle is run twice on the same inputs with a different number of BLAS threads (= a different fp32
accumulation order inside every matmul, nothing else) and once in fp32.  Result on the build box (round 2, recorded in
profiles/r02_parity_floor.md): activations differ by ~1e-2 normwise between the two bf16 runs — as much as bf16 differs
from fp32 — while the scalar losses agree to ~2e-5.  This is the floor under every activation-level tolerance in
tests/test_gpu_fullsize.py."""
import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from lap_b200 import params as P
from lap_b200.config import get_config
from lap_b200.data import synthetic_batch
from oracle import lap_oracle as O
from tests.helpers import obs_for_oracle, rel_err
tc = get_config("lap_libero"); cfg = tc.model
ref = P.init_reference_params(cfg, 7, reference_zero_init=False)
B=2
b = synthetic_batch(cfg, B, step=11); b["sample_mask"][:] = True
b["image_mask"]["left_wrist_0_rgb"][1] = False; b["image"]["left_wrist_0_rgb"][1] = -1.0
t = lambda x: torch.from_numpy(np.asarray(x))
res={}
for nt in (8, 3):
    torch.set_num_threads(nt)
    with torch.no_grad():
        t0=time.time(); loss, m, aux = O.compute_loss(ref, cfg, obs_for_oracle(b), t(b["actions"]), t(b["noise"]), t(b["time"]), bf16=True, return_aux=True)
    res[nt]=(float(loss), aux["v_t"].clone(), aux["prefix_out"].clone(), aux["suffix_out"].clone()); print(nt, float(loss), time.time()-t0, flush=True)
print("v_t rel (8 vs 3 threads):", rel_err(res[8][1], res[3][1]))
print("prefix_out rel:", rel_err(res[8][2], res[3][2]), "suffix_out rel:", rel_err(res[8][3], res[3][3]))
# fp32 mode vs bf16 mode for scale
torch.set_num_threads(8)
with torch.no_grad():
    loss32, m32, aux32 = O.compute_loss(ref, cfg, obs_for_oracle(b), t(b["actions"]), t(b["noise"]), t(b["time"]), bf16=False, return_aux=True)
print("bf16 vs fp32: v_t", rel_err(res[8][1], aux32["v_t"]), "prefix_out", rel_err(res[8][2], aux32["prefix_out"]), "suffix_out", rel_err(res[8][3], aux32["suffix_out"]), "loss", float(loss32), res[8][0])
