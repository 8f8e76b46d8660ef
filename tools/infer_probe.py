import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200.config import get_config
from lap_b200 import ops
from lap_b200.data import synthetic_batch
from lap_b200.model import LAP
from lap_b200.observation import Observation
tc = get_config("lap_libero")
model = LAP(tc.model, seed=0)
b = synthetic_batch(tc.model, 1, step=0, with_langact=False)
obs = Observation.from_dict(b)
for _ in range(3): model.sample_actions(0, obs, num_steps=10, noise=b["noise"])
torch.cuda.synchronize()
g = model._infer_graphs[(1, 10)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1)/20:.2f} ms", flush=True)
t0 = time.perf_counter()
for _ in range(20): st = model._stage(obs, with_loss=False)
torch.cuda.synchronize()
print(f"stage: {(time.perf_counter()-t0)/20*1e3:.2f} ms", flush=True)
# eager pieces
model.use_cuda_graph = False; model._infer_graphs.clear()
st = model._stage(obs, with_loss=False)
def tm(fn, n=5):
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
cfg = tc.model
import math
from lap_b200.model import _round_up
B, Pn = 1, cfg.prefix_len
X0 = model.buf("inf.X0", (B * Pn, cfg.gemma.width))
print(f"siglip fwd (eager): {tm(lambda: model._siglip_fwd(st, X0, Pn)):.2f} ms", flush=True)
Tpad = _round_up(cfg.prefix_len + cfg.action_horizon, 32)
bits_p = model._bufs["inf.bits_p"]; pos_p = model._bufs["inf.pos_p"]; Kc = model._bufs["inf.Kc"]; Vc = model._bufs["inf.Vc"]
print(f"gemma prefix (eager): {tm(lambda: model._gemma_fwd_prefix(B, X0, bits_p, pos_p, (Kc, Vc))):.2f} ms", flush=True)
gp = torch.cuda.CUDAGraph()
with torch.cuda.graph(gp):
    model._siglip_fwd(st, X0, Pn)
print(f"siglip fwd (graph): {tm(gp.replay):.2f} ms", flush=True)
gq = torch.cuda.CUDAGraph()
with torch.cuda.graph(gq):
    model._gemma_fwd_prefix(B, X0, bits_p, pos_p, (Kc, Vc))
print(f"gemma prefix (graph): {tm(gq.replay):.2f} ms", flush=True)
