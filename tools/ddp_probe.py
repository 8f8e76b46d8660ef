"""Multi-GPU phase timing: LLM fwd/bwd graph, all-reduce, SigLIP bwd graph, optimizer graph — sequential vs overlapped."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from lap_b200 import ops
from lap_b200.config import get_config
from lap_b200.data import synthetic_batch
from lap_b200.train import TrainingStepRunner, batch_from_dict, init_train_state

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
tc = get_config("lap_libero")
state = init_train_state(tc, seed=0); runner = TrainingStepRunner(tc); model = state.model
obs, actions, extra = batch_from_dict(synthetic_batch(tc.model, 32, step=0, rank=rank))
st = model._stage(obs, actions, extra["noise"], extra["time"], with_loss=True, global_counts=(58.0, 64.0))
for _ in range(4): runner.step_staged(state, st)
torch.cuda.synchronize(); dist.barrier()
g = runner._graphs[(st.B, st.R)]
lo, hi = model.llm_grad_range()
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(fn, n=3):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = ev(), ev(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def ar_all():
    runner._allreduce_end(runner._allreduce_begin(model.G))
def ar_llm():
    runner._allreduce_end(runner._allreduce_begin(model.G[lo:hi]))
def seq():
    g[0].replay(); ar_all(); g[1].replay(); g[2].replay()
def ovl():
    g[0].replay(); w = runner._allreduce_begin(model.G[lo:hi]); g[1].replay()
    w += runner._allreduce_begin(model.G[:lo]) + runner._allreduce_begin(model.G[hi:]); runner._allreduce_end(w); g[2].replay()
res = dict(g0=timeit(g[0].replay), g1=timeit(g[1].replay), g2=timeit(g[2].replay), ar_all=timeit(ar_all), ar_llm=timeit(ar_llm), seq=timeit(seq), ovl=timeit(ovl))
for bucket in (64 << 20, 2048 << 20):
    runner.bucket_bytes = bucket
    res[f"ar_all_bucket{bucket>>20}MB"] = timeit(ar_all)
if rank == 0: print({k: round(v, 1) for k, v in res.items()}, flush=True)
dist.destroy_process_group()
