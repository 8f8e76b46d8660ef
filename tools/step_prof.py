"""One LAP-3B train step with every GEMM call logged (shape, flags) so an ncu launch list can be joined by order."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200.config import get_config
from lap_b200 import ops
from lap_b200.data import synthetic_batch
from lap_b200.train import init_train_state, TrainingStepRunner, batch_from_dict

tc = get_config("lap_libero")
state = init_train_state(tc, seed=0)
runner = TrainingStepRunner(tc)
b = synthetic_batch(tc.model, 32, step=0); obs, actions, extra = batch_from_dict(b)
st = state.model._stage(obs, actions, extra["noise"], extra["time"], with_loss=True)
runner.step_staged(state, st); runner.step_staged(state, st)
torch.cuda.synchronize()
log = []
orig = ops.gemm
def logged(A, B, C, **kw):
    log.append({k: kw.get(k, d) for k, d in dict(M=0, N=0, K=0, a_major=0, b_major=0, batch_i=1, batch_o=1, epi=0).items()}
               | {"f32": C.dtype == torch.float32, "accumulate": bool(kw.get("accumulate", False)),
                  "c2": kw.get("C2") is not None})
    return orig(A, B, C, **kw)
ops.gemm = logged
import lap_b200.model as mm
torch.cuda.nvtx.range_push("STEP")
runner.step_staged(state, st)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
json.dump(log, open("gpurun_out/gemm_calls.json", "w"))
print("gemm calls", len(log))
