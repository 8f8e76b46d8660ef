"""Training runtime of the hot path: TrainState, the train step and its data-parallel gradient exchange.

Reference:
  TrainState .............. src/lap/training/state.py:8-18
  init_train_state ........ scripts/train.py:201-326
  TrainingStepRunner ...... scripts/train.py:329-419  (value_and_grad -> clip -> AdamW -> EMA -> norms)
  optimizer / schedule .... third_party/openpi/src/openpi/training/optimizer.py
  data-parallel sharding .. src/lap/training/mh_sharding.py:14-63 + implicit GSPMD all-reduce (scripts/train.py:532-537)

Design: one process per GPU; parameters, Adam moments and EMA are replicated as flat fp32 buffers (fsdp_devices=1 in
the reference); the global batch is sharded on axis 0; the ONLY data-path collective is one sum-all-reduce of the flat
fp32 gradient buffer over NCCL (bucketed so the tail of backward overlaps it).  Loss normalisers are global counts
(lap.py:580-589), exchanged as two scalars before the step.
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .config import TrainConfig
from .model import LAP, F32
from .observation import CoTObservation, Observation, to_numpy


@dataclasses.dataclass
class TrainState:
    """src/lap/training/state.py:8-18 — engine-owned flat device buffers (donated / updated in place)."""

    step: int
    model: LAP  # owns params (fp32 master `P`) and the bf16 compute copy
    mu: torch.Tensor
    nu: torch.Tensor
    ema_params: torch.Tensor | None
    ema_decay: float | None

    @property
    def params(self) -> torch.Tensor:
        return self.model.P


def init_train_state(config: TrainConfig, seed: int | None = None, *, model: LAP | None = None,
                     reference_zero_init: bool = True) -> TrainState:
    """scripts/train.py:201-326 with weight_loader.kind=none (random init) unless a model is passed in."""
    if model is None:
        model = LAP(config.model, seed=config.seed if seed is None else seed, reference_zero_init=reference_zero_init)
    n = model.layout.total
    dev = model.device
    model.G = torch.zeros(n, dtype=F32, device=dev)
    ema_decay, ema_enabled = config.get_ema_init()
    return TrainState(step=0, model=model, mu=torch.zeros(n, dtype=F32, device=dev),
                      nu=torch.zeros(n, dtype=F32, device=dev),
                      ema_params=model.P.clone() if ema_enabled else None, ema_decay=ema_decay)


class TrainingStepRunner:
    """Callable with the reference's signature: (rng, state, (observation, actions), step) -> (state, info)."""

    def __init__(self, config: TrainConfig, *, bucket_bytes: int = 512 << 20, use_cuda_graph: bool = True):
        self.config = config
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.bucket_bytes = bucket_bytes
        self.use_cuda_graph = use_cuda_graph
        # Defaults = the measured best at N = 2 and N = 8 (profiles/r02_overlap_sweep.md): ONE LLM group + ONE SigLIP group
        # (the LLM gradients are reduced under the SigLIP backward) and 16 SMs left to NCCL (NVLS runs 16 CTAs whatever
        # NCCL_MAX_NCHANNELS says).  Finer groups hide more of the all-reduce but put the SM carve-out on more GEMMs: equal
        # within noise at N = 8 (373.9 vs 374.9 ms), slower at N = 2.
        self.comm_sms = int(os.environ.get("LAPB_COMM_SMS", "16"))  # SMs left to NCCL while it overlaps compute
        self.bwd_segments = int(os.environ.get("LAPB_BWD_SEGMENTS", "1"))  # LLM backward groups (world > 1)
        self.vis_segments = int(os.environ.get("LAPB_VIS_SEGMENTS", "1"))  # SigLIP backward groups (world > 1)
        self._phase_cache: dict = {}
        self._partials = None
        self._stats = None
        self._hyper_host = None
        self._graphs: dict = {}  # (B, R_cap, state, model) -> (one graph per phase ..., optimizer graph)
        self._warm: dict = {}  # eager steps run per (B, R_cap): workspaces exist before capture

    def reset(self) -> None:
        """Forget the captured step graphs (call before `LAP.release_workspaces`, or after replacing the train state)."""
        torch.cuda.synchronize()
        self._graphs.clear()
        self._warm.clear()
        self._phase_cache.clear()

    # -- global loss normalisers (lap.py:580-589 are means over the GLOBAL batch) ------------------------------
    def _global_counts(self, observation, B: int, device) -> tuple[float, float]:
        sm = getattr(observation, "sample_mask", None)
        n_active = float(to_numpy(sm).astype(bool).sum()) if sm is not None else float(B)
        n_action = float(B)
        if self.world > 1:
            t = torch.tensor([n_active, n_action], dtype=torch.float32, device=device)  # counts <= 2^24: exact in fp32
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            n_active, n_action = (float(x) for x in t.tolist())
        return n_active, n_action

    def _allreduce_begin(self, G: torch.Tensor) -> list:
        """C1: the one data-path collective — sum of fp32 grads across ranks (NCCL over NVLink/NVSwitch), issued
        asynchronously in buckets; the collective runs on NCCL's stream and overlaps whatever is enqueued next."""
        if self.world == 1 or G.numel() == 0:
            return []
        step = self.bucket_bytes // 4
        return [dist.all_reduce(G[o : o + step], op=dist.ReduceOp.SUM, async_op=True)
                for o in range(0, G.numel(), step)]

    @staticmethod
    def _allreduce_end(works: list) -> None:
        for w in works:
            w.wait()

    def _allreduce_grads(self, G: torch.Tensor) -> None:
        self._allreduce_end(self._allreduce_begin(G))

    def __call__(self, rng, state: TrainState, batch, step: int | None = None, *, with_metrics: bool = True):
        cfg = self.config
        model = state.model
        observation, actions = batch[0], batch[1]
        extra = batch[2] if len(batch) > 2 else {}
        noise, time = extra.get("noise"), extra.get("time")
        step = state.step if step is None else int(step)
        B = to_numpy(actions).shape[0] if not isinstance(actions, torch.Tensor) else actions.shape[0]
        mc = cfg.model
        if noise is None or time is None:
            # train_rng = fold_in(rng, step) (scripts/train.py:355); torch generator stands in for threefry
            # every data-parallel rank draws for ITS shard of the global batch: the rank is folded into the seed (the
            # reference draws independent noise / time for each sample of the global batch)
            rank = dist.get_rank() if self.world > 1 else 0
            gen = torch.Generator().manual_seed(((int(rng) if rng is not None else 0) * 1_000_003 + step) * 4099 + rank)
            if noise is None:
                noise = torch.randn((B, mc.action_horizon, mc.action_dim), generator=gen)
            if time is None:
                u = torch.rand((B,), generator=gen)
                time = u.pow(1.0 / 1.5) * 0.999 + 0.001  # Beta(1.5, 1) by inverse CDF
        aug = extra.get("aug")
        if aug is None and mc.enable_image_augmentation:
            # model_adapter.py:118-151 is on for training unless the config disables it (lap_libero does); the draw is
            # per rank and step like the flow-matching noise
            rank = dist.get_rank() if self.world > 1 else 0
            aug = model.draw_augmentation(((int(rng) if rng is not None else 0) * 1_000_003 + step) * 4099 + rank,
                                          observation, B)
        counts = self._global_counts(observation, B, model.device)
        st = model._stage(observation, actions, noise, time, with_loss=True, global_counts=counts, aug=aug)
        info = self.step_staged(state, st, step)
        if self.world > 1:
            # each rank holds its shard's share of the global mean; the sum over ranks is the global loss
            lt = info["loss"].clone()
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            info["loss"] = lt
        if with_metrics:
            info.update(model._metrics(st))
        return state, info

    def _phases(self, model: LAP):
        """The step as a list of (compute, gradient ranges final after it, overlaps-communication) phases.

        world == 1: [forward + whole LLM backward], [SigLIP backward].
        world > 1 : the LLM backward is cut into `bwd_segments` groups of layers and the SigLIP backward into
        `vis_segments`, and each group's weight-gradient ranges are all-reduced while the NEXT group computes
        (scripts/train.py:532-537: the reference's all-reduce is scheduled by XLA inside the backward the same way).
        Only the small tail (biases, norm scales, position table) and the last SigLIP group are reduced with nothing
        left to hide them."""
        lay = model.layout
        L, Ls = model.cfg.gemma.depth, model.cfg.siglip.depth
        nl = max(1, min(self.bwd_segments if self.world > 1 else 1, L))
        nv = max(1, min(self.vis_segments if self.world > 1 else 1, Ls))
        lcuts = [round(L * i / nl) for i in range(nl, -1, -1)]      # e.g. [18, 12, 6, 0]
        vcuts = [round(Ls * i / nv) for i in range(nv, -1, -1)]     # e.g. [27, 18, 9, 0]
        llm_rest = (lay.offsets["e.mod_w"], lay.small_begin)        # modulation Dense, action / time projections, embedding
        phases = []
        for i in range(nl):
            hi, lo = lcuts[i], lcuts[i + 1]

            def run(st, hi=hi, lo=lo, first=(i == 0), last=(i == nl - 1)):
                if first:
                    model.forward_and_heads(st)
                model.backward_llm_layers(st, hi, lo)
                if last:
                    model.backward_llm_tail(st)

            ranges = model.layer_grad_ranges(model.LLM_LAYER_GRADS, lo, hi) + ([llm_rest] if i == nl - 1 else [])
            phases.append((run, ranges, i > 0))
        for i in range(nv):
            hi, lo = vcuts[i], vcuts[i + 1]

            def runv(st, hi=hi, lo=lo):
                model.backward_vision_segment(st, hi, lo)

            ranges = model.layer_grad_ranges(model.VIS_LAYER_GRADS, lo, hi)
            if i == 0:
                ranges.append((lay.offsets["img.head_w"], lay.offsets["g.qkv_w"]))
            if i == nv - 1:
                ranges += [(lay.offsets["img.patch_w"], lay.offsets["img.qkv_w"]), (lay.small_begin, lay.total)]
            phases.append((runv, ranges, True))
        return phases

    def _run_phase(self, fn, st, overlaps_comm: bool) -> None:
        """With world > 1 the persistent GEMMs of a phase that runs next to an in-flight all-reduce leave `comm_sms`
        SMs to NCCL (a 148-CTA persistent grid on fewer free SMs would run a second wave)."""
        if self.world > 1 and overlaps_comm:
            ops.gemm_max_ctas = max(2, (ops.num_sms() - self.comm_sms) // 2 * 2)
        try:
            fn(st)
        finally:
            ops.gemm_max_ctas = 0

    def step_staged(self, state: TrainState, st, step: int | None = None) -> dict:
        """One optimisation step on inputs already staged in HBM (model._stage) — no host<->device traffic except
        the 32-byte per-step hyper-parameter vector.  After two eager steps (which allocate every workspace) the
        phases (`_phases`) and the optimizer are captured as CUDA graphs and replayed: ~1700 kernel launches become a
        handful of graph launches.  With world > 1 each phase's finished gradient ranges are all-reduced (NCCL,
        asynchronously) while the next phase computes."""
        model = state.model
        step = state.step if step is None else int(step)
        self._set_hyper(state, step)
        key = (st.B, st.R, id(state), id(model))  # a graph is bound to the buffers of ONE state / model
        phases = self._phase_cache.get(id(model))
        if phases is None:
            phases = self._phase_cache[id(model)] = self._phases(model)
        g = self._graphs.get(key)
        if g is None and self.use_cuda_graph and self._warm.get(key, 0) >= 2:
            g = self._capture(state, st, phases)
            self._graphs[key] = g
        works = []
        for i, (fn, ranges, overlaps) in enumerate(phases):
            if g is None:
                self._run_phase(fn, st, overlaps)
            else:
                g[i].replay()
            for lo, hi in ranges:
                works += self._allreduce_begin(model.G[lo:hi])
        self._allreduce_end(works)
        if g is None:
            self._apply_gradients_device(state)
            self._warm[key] = self._warm.get(key, 0) + 1
        else:
            g[-1].replay()
        loss = model.buf("loss.total", (1,), F32)
        state.step = step + 1
        # fresh scalars (the reference returns new arrays): the persistent buffers are overwritten by the next step,
        # and callers accumulate info dicts over log_interval steps.  Stream-ordered clones, no host sync.
        stats = self._stats.clone()
        return {"loss": loss[0].clone(), "grad_norm": stats[0], "grad_norm_f32": stats[0], "param_norm": stats[3]}

    def _capture(self, state: TrainState, st, phases):
        model = state.model
        torch.cuda.synchronize()
        pool = torch.cuda.graph_pool_handle()
        graphs = []
        for fn, _, overlaps in phases:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self._run_phase(fn, st, overlaps)
            graphs.append(g)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            self._apply_gradients_device(state)
        graphs.append(g)
        torch.cuda.synchronize()
        # capture does not execute: the caller replays the graphs right away to perform this step
        return tuple(graphs)

    def _set_hyper(self, state: TrainState, step: int) -> None:
        """Per-step scalars (lr, Adam bias corrections, EMA decay) go to the device as one tiny pinned copy."""
        cfg, o = self.config, self.config.optimizer
        model = state.model
        if self._partials is None:
            self._np = ops.opt_num_partials()
            self._partials = torch.zeros(self._np, dtype=F32, device=model.device)
            self._stats = torch.zeros(4, dtype=F32, device=model.device)
            # ring of pinned staging vectors: slot i is rewritten only after the copy that last read it has executed
            self._hyper_host = [torch.zeros(8, dtype=F32).pin_memory() for _ in range(4)]
            self._hyper_events = [None] * 4
            self._hyper_slot = 0
            self._hyper = torch.zeros(8, dtype=F32, device=model.device)
        count = step + 1
        decay, ema_on = cfg.get_ema_decay_for_step(step)
        slot = self._hyper_slot
        self._hyper_slot = (slot + 1) % len(self._hyper_host)
        if self._hyper_events[slot] is not None:
            self._hyper_events[slot].synchronize()
        h = self._hyper_host[slot]
        h[0] = cfg.lr_schedule.lr(step)
        h[1] = 1.0 - o.b1 ** count
        h[2] = 1.0 - o.b2 ** count
        h[3] = float(decay)
        h[4] = 1.0 if (ema_on and state.ema_params is not None) else 0.0
        self._hyper.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._hyper_events[slot] = ev

    def share_scratch(self, other: "TrainingStepRunner") -> None:
        """Use another runner's optimizer scratch (norm partials, stats, hyper-parameter staging): bench.py replays the
        timed steps through a second, eager runner to time individual launches."""
        for n in ("_partials", "_stats", "_hyper_host", "_hyper_events", "_hyper_slot", "_hyper", "_np"):
            setattr(self, n, getattr(other, n))

    def _apply_gradients_device(self, state: TrainState) -> None:
        """clip_by_global_norm -> adamw -> apply -> EMA (scripts/train.py:363-396) + norms (:370-371,402-415); all
        step-dependent scalars are read from device memory (self._hyper)."""
        o = self.config.optimizer
        model = state.model
        n = model.layout.total
        self._stats[2:].zero_()
        ops.sumsq_partials(model.G, n, self._partials)
        ops.adamw_ema(model.P, model.G, state.mu, state.nu, state.ema_params, model.W16, n, self._partials, self._np,
                      self._stats, 0, model.layout.kernel_end, lr=0.0, b1=o.b1, b2=o.b2, eps=o.eps, wd=o.weight_decay,
                      bc1=1.0, bc2=1.0, clip=o.clip_gradient_norm, ema_decay=0.0, ema_on=False, hyper=self._hyper)
        ops.sqrt_scalar(self._stats, 2, 3)
        model.refresh_embed_split()

    def apply_gradients(self, state: TrainState, step: int) -> dict:
        """Eager optimizer step for callers that produced model.G themselves."""
        self._set_hyper(state, step)
        self._apply_gradients_device(state)
        state.step = step + 1
        stats = self._stats.clone()
        return {"grad_norm": stats[0], "grad_norm_f32": stats[0], "param_norm": stats[3]}


class ValidationStepRunner:
    """scripts/train.py:422-450: `compute_loss(train=False)` on a validation batch -> the metrics dict plus `val_loss`.
    The reference folds `state.step` into the rng (`fold_in(rng, state.step)`); explicit `noise` / `time` in the batch
    extras take precedence, as in `TrainingStepRunner`."""

    def __init__(self, config: TrainConfig):
        self.config = config

    def __call__(self, rng, state: TrainState, batch) -> dict:
        observation, actions = batch[0], batch[1]
        extra = batch[2] if len(batch) > 2 else {}
        eval_rng = (int(rng) if rng is not None else 0) * 1_000_003 + int(state.step)
        val_loss, val_metrics = state.model.compute_loss(
            eval_rng, observation, actions, train=False, verbose_mode=self.config.model.verbose_mode,
            noise=extra.get("noise"), time=extra.get("time"))
        val_metrics = dict(val_metrics)
        val_metrics["val_loss"] = val_loss
        return val_metrics


def train_step(config: TrainConfig, rng, state: TrainState, batch, step: int | None = None):
    return TrainingStepRunner(config)(rng, state, batch, step)


def batch_from_dict(d: dict):
    """Loader-format dict (data_loader.py:327) -> ((CoTObservation, actions), extras)."""
    obs = CoTObservation.from_dict(d)
    return obs, d["actions"], {k: d[k] for k in ("noise", "time", "aug") if k in d}
