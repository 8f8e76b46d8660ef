"""The caller of the train step: the step loop of `scripts/train.py:main` (:458-620) around `TrainingStepRunner`, with the
reference's resume / save / log cadence.  Data loading (RLDS / tf.data), wandb and validation are outside the hot path: the
loop takes any iterable of `(observation, actions[, extras])` batches in the loader format and a logging callable.

    state = run_training(get_config("lap_libero"), batches, checkpoint_dir="ckpt/lap_libero/run0")

One process per GPU (`torchrun`): every rank runs the loop on its shard of the global batch; rank 0 writes checkpoints.
"""
from __future__ import annotations

from collections.abc import Callable, Iterable

import torch

from . import checkpoint as _checkpoints
from .config import TrainConfig


def _mean_infos(infos: list[dict]) -> dict:
    """metrics_logging.process_and_log_metrics reduces the infos gathered since the last log by their mean."""
    out = {}
    for k in infos[0]:
        vals = [float(i[k]) for i in infos if k in i and i[k] is not None and getattr(i[k], "numel", lambda: 1)() == 1]
        if vals:
            out[k] = sum(vals) / len(vals)
    return out


def run_training(config: TrainConfig, batches: Iterable, *, checkpoint_dir=None, resume: bool = True, state=None,
                 runner=None, log_fn: Callable[[int, dict], None] | None = None, rank: int | None = None):
    """Train from `state.step` (after an optional resume) to `config.num_train_steps`.  Returns the final TrainState.

    * resume (scripts/train.py:468-530): with `resume` and a checkpoint under `checkpoint_dir`, the newest one is restored and
      the loop continues from its step;
    * save (:586-598): `(step % save_interval == 0 and step > start_step) or step == num_train_steps`, newest checkpoint kept
      plus every `keep_period`-th step (checkpoints.py:58-63);
    * log (:600-617): every `log_interval` steps, the mean of the step infos since the last log.
    """
    if state is None or runner is None:
        from .train import TrainingStepRunner, init_train_state  # needs the CUDA engine
        state = init_train_state(config) if state is None else state
        runner = TrainingStepRunner(config) if runner is None else runner
    if rank is None:
        rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
    if checkpoint_dir is not None and resume and _checkpoints.latest_step(checkpoint_dir) is not None:
        _checkpoints.restore_train_state(checkpoint_dir, state)
    start_step = int(state.step)
    it = iter(batches)
    infos: list[dict] = []
    for step in range(start_step, config.num_train_steps):
        try:
            batch = next(it)
        except StopIteration:
            break
        state, info = runner(config.seed, state, batch, step)
        infos.append(info)
        should_save = (step % config.save_interval == 0 and step > start_step) or step == config.num_train_steps
        if should_save and checkpoint_dir is not None and rank == 0:
            _checkpoints.save_train_state(checkpoint_dir, state, step, keep=1, keep_period=config.keep_period)
        if step % config.log_interval == 0:
            if log_fn is not None and rank == 0:
                log_fn(step, _mean_infos(infos))
            infos = []
    return state
