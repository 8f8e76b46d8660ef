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
