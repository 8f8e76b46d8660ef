"""Full-size LAP-3B probe: memory + step time + phase breakdown (CUDA events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200.config import get_config
from lap_b200 import ops
from lap_b200.data import synthetic_batch
from lap_b200.train import init_train_state, TrainingStepRunner, batch_from_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tc = get_config("lap_libero")
t0 = time.time()
state = init_train_state(tc, seed=0)
torch.cuda.synchronize()
print(f"init {time.time()-t0:.1f}s  mem {torch.cuda.memory_allocated()/2**30:.1f} GiB", flush=True)
runner = TrainingStepRunner(tc)
model = state.model
for i in range(steps + 3):
    b = synthetic_batch(tc.model, B, step=i)
    obs, actions, extra = batch_from_dict(b)
    torch.cuda.synchronize(); t0 = time.time()
    n0 = ops.launch_count
    state, info = runner(0, state, (obs, actions, extra), with_metrics=False)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"step {i}: {dt*1e3:.1f} ms  {B/dt:.2f} samples/s  loss {float(info['loss']):.4f} gnorm {float(info['grad_norm']):.4f} launches {ops.launch_count-n0}  mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
# phase breakdown
b = synthetic_batch(tc.model, B, step=99); obs, actions, extra = batch_from_dict(b)
ev = lambda: torch.cuda.Event(enable_timing=True)
st = model._stage(obs, actions, extra["noise"], extra["time"], with_loss=True)
e = [ev() for _ in range(6)]
torch.cuda.synchronize()
e[0].record(); loss, _ = model._forward_loss(st, save=True, compute_grad_seed=True); e[1].record()
torch.cuda.synchronize()
print(f"forward+loss: {e[0].elapsed_time(e[1]):.1f} ms", flush=True)
e[2].record(); model.forward_backward(st); e[3].record(); torch.cuda.synchronize()
print(f"forward+backward: {e[2].elapsed_time(e[3]):.1f} ms", flush=True)
e[4].record(); runner.apply_gradients(state, 5); e[5].record(); torch.cuda.synchronize()
for _ in range(3): runner.step_staged(state, st)
torch.cuda.synchronize()
e[0].record()
for _ in range(5): runner.step_staged(state, st)
e[1].record(); torch.cuda.synchronize()
print(f"graphed step_staged: {e[0].elapsed_time(e[1])/5:.1f} ms/step", flush=True)
print(f"optimizer: {e[4].elapsed_time(e[5]):.1f} ms", flush=True)
# host-side time of a step (python overhead)
t0 = time.time(); model.forward_backward(st); t1 = time.time(); torch.cuda.synchronize(); t2 = time.time()
print(f"host enqueue time {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms")
print("DONE")
