import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200.config import get_config
from lap_b200.data import synthetic_batch
from lap_b200.train import init_train_state, TrainingStepRunner, batch_from_dict
tc = get_config("lap_libero")
state = init_train_state(tc, seed=0)
rg = TrainingStepRunner(tc, use_cuda_graph=True); re_ = TrainingStepRunner(tc, use_cuda_graph=False)
model = state.model
obs, actions, extra = batch_from_dict(synthetic_batch(tc.model, 32, step=0))
st = model._stage(obs, actions, extra["noise"], extra["time"], with_loss=True)
for _ in range(4): rg.step_staged(state, st)
re_.share_scratch(rg)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def tm(r, n=10):
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): r.step_staged(state, st)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for rnd in range(3):
    print(f"round {rnd}: graph {tm(rg):.1f} ms   eager {tm(re_):.1f} ms", flush=True)
g = rg._graphs[(st.B, st.R)]
def tg(i, n=10):
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): g[i].replay()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print("graphs:", [round(tg(i), 1) for i in range(3)])
