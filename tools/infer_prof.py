import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200.config import get_config
from lap_b200 import ops
from lap_b200.data import synthetic_batch
from lap_b200.model import LAP
from lap_b200.observation import Observation
tc = get_config("lap_libero")
model = LAP(tc.model, seed=0)
model.use_cuda_graph = False
b = synthetic_batch(tc.model, 1, step=0, with_langact=False)
obs = Observation.from_dict(b)
for _ in range(2): model.sample_actions(0, obs, num_steps=10, noise=b["noise"])
torch.cuda.synchronize()
log = []
orig = ops.gemm
def logged(A, B, C, **kw):
    log.append({k: kw.get(k, d) for k, d in dict(M=0, N=0, K=0, a_major=0, b_major=0, batch_i=1, batch_o=1, epi=0).items()} | {"f32": C.dtype == torch.float32})
    return orig(A, B, C, **kw)
ops.gemm = logged
torch.cuda.nvtx.range_push("STEP")
model.sample_actions(0, obs, num_steps=10, noise=b["noise"])
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
json.dump(log, open("gpurun_out/gemm_calls_infer.json", "w"))
print("gemm calls", len(log), "launches", ops.launch_count)
