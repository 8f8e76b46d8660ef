"""world_size-2 gloo test (CPU) of the data-parallel logic: sharding the global batch with GLOBAL loss normalisers and
summing the shard gradients reproduces the full-batch gradient, using the same host code the GPU path
uses for the count exchange and the bucketed all-reduce."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lap_b200 import params as P
    from lap_b200.config import get_config
    from lap_b200.data import synthetic_batch
    from lap_b200.observation import CoTObservation
    from lap_b200.train import TrainingStepRunner
    from oracle import lap_oracle as O
    from tests.helpers import obs_for_oracle

    torch.set_num_threads(2)
    tc = get_config("debug_tiny")
    cfg = tc.model
    ref = P.init_reference_params(cfg, 0, reference_zero_init=False)
    full = synthetic_batch(cfg, 4, step=5)
    full["sample_mask"] = np.array([True, False, True, True])
    sl = slice(rank * 2, rank * 2 + 2)
    shard = {k: ({kk: vv[sl] for kk, vv in v.items()} if isinstance(v, dict) else v[sl]) for k, v in full.items()}
    runner = TrainingStepRunner(tc, bucket_bytes=4096)
    assert runner.world == 2
    n_active, n_action = runner._global_counts(CoTObservation.from_dict(shard), 2, torch.device("cpu"))
    assert (n_active, n_action) == (3.0, 4.0)
    t = lambda x: torch.from_numpy(np.asarray(x))
    # shard loss with GLOBAL normalisers: sum_shard(lang_w)/n_active + sum_shard(act_w)/n_action
    params = {k: v.clone().requires_grad_(True) for k, v in ref.items()}
    _, _, aux = O.compute_loss(params, cfg, obs_for_oracle(shard), t(shard["actions"]), t(shard["noise"]), t(shard["time"]), bf16=False, return_aux=True)
    loss = cfg.language_loss_weight * aux["lang_per_sample"].sum() / n_active + cfg.action_loss_weight * aux["action_per_sample"].sum() / n_action
    loss.backward()
    flat = torch.cat([params[k].grad.reshape(-1) for k in sorted(params)])
    runner._allreduce_grads(flat)  # bucketed sum all-reduce (gloo here, NCCL on the GPUs)
    lt = loss.detach().clone()
    dist.all_reduce(lt)
    if rank == 0:
        pf = {k: v.clone().requires_grad_(True) for k, v in ref.items()}
        lf, _ = O.compute_loss(pf, cfg, obs_for_oracle(full), t(full["actions"]), t(full["noise"]), t(full["time"]), bf16=False)
        lf.backward()
        flat_full = torch.cat([pf[k].grad.reshape(-1) for k in sorted(pf)])
        q.put((float(lt), float(lf), float((flat - flat_full).norm() / flat_full.norm())))
    dist.destroy_process_group()


def test_sharded_gradients_sum_to_global_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    loss_sum, loss_full, rel = q.get(timeout=10)
    assert abs(loss_sum - loss_full) < 1e-5 * abs(loss_full)
    assert rel < 1e-4


def _loop_worker(rank, world, port, ckpt_dir, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dataclasses
    from lap_b200 import checkpoint as C, params as P
    from lap_b200.config import get_config
    from lap_b200.train import TrainState
    from lap_b200.train_loop import run_training

    cfg = dataclasses.replace(get_config("debug_tiny"), num_train_steps=9, save_interval=3, keep_period=None, log_interval=100)

    class CpuModel:
        def __init__(self, mc):
            self.cfg, self.layout = mc, P.FlatLayout(mc)
            self.P = torch.zeros(self.layout.total)
        def params_reference(self, flat=None):
            eng = {k: v.detach().float().cpu() for k, v in P.engine_from_flat(self.layout, self.P if flat is None else flat).items()}
            return P.engine_to_reference(self.cfg, eng)
        def refresh_compute_copy(self):
            pass

    m = CpuModel(cfg.model)
    n = m.layout.total
    state = TrainState(step=0, model=m, mu=torch.zeros(n), nu=torch.zeros(n), ema_params=None, ema_decay=None)
    saves = []
    real_save = C.save_train_state
    C.save_train_state = lambda *a, **k: (saves.append(a[2]), real_save(*a, **k))[1]

    def runner(rng, st, batch, step):            # replicated "training": every rank applies the all-reduced update
        g = torch.tensor([float(batch)])
        dist.all_reduce(g)
        st.model.P += g
        st.step = step + 1
        return st, {"loss": g[0]}

    state = run_training(cfg, iter(range(rank, 100, world)), checkpoint_dir=ckpt_dir, state=state, runner=runner)
    dist.barrier()
    q.put((rank, saves, float(state.model.P[0]), sorted(int(p) for p in os.listdir(ckpt_dir) if p.isdigit())))
    dist.destroy_process_group()


def test_training_loop_saves_on_rank0_only(tmp_path):
    """`run_training` under world_size 2 (gloo): the ranks stay in lock-step through the collective in the step, only rank 0
    writes checkpoints (the state is replicated), and both end with the same parameters."""
    import pytest
    pytest.importorskip("safetensors")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + os.getpid() % 1000
    procs = [ctx.Process(target=_loop_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    got = dict((r, rest) for r, *rest in (q.get(timeout=10) for _ in range(2)))
    assert got[0][0] == [3, 6] and got[1][0] == []            # rank 0 saved at steps 3 and 6, rank 1 never
    assert got[0][1] == got[1][1] == float(sum(range(18)))    # 9 steps x 2 shards, all-reduced
    assert got[0][2] == [6]                                    # max_to_keep = 1
