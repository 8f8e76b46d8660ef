#!/usr/bin/env python
"""bench.py — LAP-3B training throughput on B200 (the BASELINE.json metric), one JSON line on stdout.

  python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores
  python bench.py --mode infer                             # action-chunk inference latency (second headline metric)

A "step" = one optimisation step (forward + backward + grad all-reduce + clip/AdamW/EMA) of the lap_libero
configuration (LAP-3B: SigLIP-So400m + Gemma-2B + Gemma-300M expert; 2x224x224x3 images, 180 text tokens, 10-step
action chunk) on a synthetic RLDS-shaped batch of 32 samples per GPU with random-init weights.
  value : samples/s with the step's inputs already resident in HBM (CUDA events, max over ranks)
  e2e   : samples/s through the public API `TrainingStepRunner(rng, state, (obs, actions))` with HOST numpy
          buffers (uint8 camera frames, the loader's wire format): pinned H2D copy of the batch and D2H read of the
          loss inside the timed region
  roofline : the tcgen05 GEMM family — every GEMM launch of the step (>90 % of step FLOPs): algorithmic FLOPs /
          summed CUDA-event time of those launches, against the measured sustained bf16 peak; per-shape table as extra
  infer : (N = 1) the second headline metric, action-chunk latency p50/p90 of `sample_actions` at batch 1, with the
          device-side split and the HBM roofline of the denoise-loop kernel
  cpu_baseline : the oracle (CPU restatement of the reference, torch fp32 eager) timed on this box's host cores on a
          bounded sample (batch 1; forward + backward + clip/AdamW/EMA, 3 timed steps) — a reported baseline, not the
          target
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LAP-3B train samples/sec"
UNIT = "samples/s"
PER_GPU_BATCH = 32


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.proc = None
        self.lines: list[str] = []
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference(steps: int, warmup: int, *, budget_s: float = 150.0) -> dict:
    """The oracle (CPU restatement of the reference) doing whole train steps of the lap_libero workload at batch 1 on
    the host cores: forward + backward (torch autograd through the fp32 restatement) + clip_by_global_norm + AdamW +
    EMA, the last three in place when the host has the memory for params + grads + mu + nu + ema (5 x 13.4 GB)."""
    import numpy as np
    import torch

    from lap_b200 import params as P
    from lap_b200.config import get_config
    from lap_b200.data import synthetic_batch
    from oracle import lap_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tc = get_config("lap_libero")
    cfg = tc.model
    t0 = time.time()
    shapes = P.reference_shapes(cfg)
    gen = torch.Generator().manual_seed(0)
    params = {}
    for k, s in shapes.items():
        t = torch.empty(s, dtype=torch.float32)
        leaf = k.rsplit("/", 1)[-1]
        if leaf == "scale" and "/img/" in k:
            t.fill_(1.0)
        elif leaf in ("bias", "scale"):
            t.zero_()
        else:
            fan_in = s[-2] if len(s) >= 2 else s[-1]
            t.normal_(0.0, 0.01 if k.endswith("input_embedding") else 1.0 / (fan_in ** 0.5), generator=gen)
        params[k] = t
    n_bytes = 4 * sum(v.numel() for v in params.values())
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    with_opt = avail > 4.6 * n_bytes + (16 << 30)  # grads + mu + nu + ema + activations next to the params
    if with_opt:
        mu = {k: torch.zeros_like(v) for k, v in params.items()}
        nu = {k: torch.zeros_like(v) for k, v in params.items()}
        ema = {k: v.clone() for k, v in params.items()}
    init_s = time.time() - t0
    tt = lambda x: torch.from_numpy(np.asarray(x))
    o = tc.optimizer

    def one_step(i: int) -> float:
        b = synthetic_batch(cfg, 1, step=i)
        b["sample_mask"][:] = True
        obs = dict(images={k: tt(v) for k, v in b["image"].items()}, image_masks={k: tt(v) for k, v in b["image_mask"].items()},
                   tokenized_prompt=tt(b["tokenized_prompt"]), tokenized_prompt_mask=tt(b["tokenized_prompt_mask"]),
                   tokenized_langact_mask=tt(b["tokenized_langact_mask"]), token_loss_mask=tt(b["token_loss_mask"]),
                   sample_mask=tt(b["sample_mask"]))
        ps = {k: v.requires_grad_(True) for k, v in params.items()}
        t1 = time.time()
        loss, _ = O.compute_loss(ps, cfg, obs, tt(b["actions"]), tt(b["noise"]), tt(b["time"]), bf16=False)
        loss.backward()
        if with_opt:  # scripts/train.py:363-396 (optax clip_by_global_norm -> adamw -> EMA), in place
            with torch.no_grad():
                gn = float(torch.sqrt(sum((v.grad.double() ** 2).sum() for v in ps.values() if v.grad is not None)))
                scale = 1.0 if gn < o.clip_gradient_norm else o.clip_gradient_norm / gn
                lr = tc.lr_schedule.lr(i)
                bc1, bc2 = 1.0 - o.b1 ** (i + 1), 1.0 - o.b2 ** (i + 1)
                decay, ema_on = tc.get_ema_decay_for_step(i)
                for k, v in ps.items():
                    if v.grad is None:
                        continue
                    g = v.grad.mul_(scale)
                    mu[k].mul_(o.b1).add_(g, alpha=1 - o.b1)
                    nu[k].mul_(o.b2).addcmul_(g, g, value=1 - o.b2)
                    upd = (nu[k] / bc2).sqrt_().add_(o.eps).reciprocal_().mul_(mu[k]).div_(bc1).add_(v, alpha=o.weight_decay)
                    v.sub_(upd, alpha=lr)
                    if ema_on:
                        ema[k].mul_(decay).add_(v, alpha=1 - decay)
        for v in ps.values():
            v.grad = None
        return time.time() - t1

    times = []
    spent = 0.0
    warmup = max(1, min(warmup, 1))  # one untimed step (first-touch of 67 GB of state); CPU steps take ~30 s each
    for i in range(warmup + steps):
        dt = one_step(i)
        spent += dt
        if i >= warmup:
            times.append(dt)
        if spent > budget_s and len(times) >= min(3, steps):
            break
    ms = 1e3 * sum(times) / len(times)
    what = "forward + backward + clip + AdamW + EMA" if with_opt else \
        "forward + backward only (host memory too small for the optimizer state)"
    return {"value": 1e3 / ms, "ms_per_step": ms, "cores": cores, "steps_timed": len(times), "init_s": init_s,
            "with_optimizer": with_opt,
            "sample": "batch 1 of the lap_libero workload, full LAP-3B (fp32 eager torch restatement of the JAX "
                      "reference: %s), %d timed steps after %d warm-up" % (what, len(times), warmup)}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps_timed"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "lap_libero LAP-3B train step, per-GPU batch 32 (reference arm: CPU, batch 1 sample)",
                   "per_gpu_batch": PER_GPU_BATCH, "images": "2x224x224x3", "text_tokens": 180, "action_horizon": 10},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "JAX/Flax are not installable here (no wheels, no network): this is the oracle port of the reference "
                "arithmetic on the host cores, not the JAX program",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_train(args) -> None:
    import torch
    import torch.distributed as dist

    from lap_b200 import ops
    from lap_b200.config import get_config
    from lap_b200.data import synthetic_batch
    from lap_b200.train import TrainingStepRunner, batch_from_dict, init_train_state

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the all-reduce of the LLM gradients overlaps the SigLIP backward: 16 NCCL channels (N = 2: the p2p ring needs them;
        # N = 8: NVLS runs 16 CTAs regardless) next to the 16 SMs the persistent GEMMs leave free
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "16")
        os.environ.setdefault("LAPB_COMM_SMS", "16")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    tc = get_config("lap_libero")
    B = PER_GPU_BATCH
    state = init_train_state(tc, seed=0)
    runner = TrainingStepRunner(tc)
    model = state.model

    def host_batch(i, uint8=False):
        return batch_from_dict(synthetic_batch(tc.model, B, step=i, rank=rank, uint8_images=uint8))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident timing (`value`) ----------------
    obs, actions, extra = host_batch(0)
    counts = runner._global_counts(obs, B, dev)
    st = model._stage(obs, actions, extra["noise"], extra["time"], with_loss=True, global_counts=counts)
    for i in range(args.warmup):  # first two are eager (allocate workspaces), the third captures the CUDA graphs
        runner.step_staged(state, st)
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        runner.step_staged(state, st)
    e1.record()
    sync()
    ms_dev = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    # roofline of the dominant kernel: the same steps once more, launched eagerly so every tcgen05 GEMM launch can be
    # bracketed by CUDA events on its stream (a captured graph cannot be instrumented per kernel)
    eager = TrainingStepRunner(tc, use_cuda_graph=False)
    eager.share_scratch(runner)
    ops.gemm_profile_begin()
    n0 = ops.launch_count
    nprof = max(1, min(args.steps, 3))
    for i in range(nprof):
        eager.step_staged(state, st)
    launches = (ops.launch_count - n0) // nprof
    gemm_flops, gemm_ms, gemm_launches, by_shape = ops.gemm_profile_end()
    gemm_flops, gemm_ms, gemm_launches = gemm_flops / nprof, gemm_ms / nprof, gemm_launches / nprof
    # ---------------- end-to-end timing through the public API with host buffers (`e2e`) ----------------
    # the loader contract ships uint8 images (data_loader.py:324): 9.6 MB per 32-sample batch, converted to
    # [-1, 1] inside the patchify kernel
    hb = [host_batch(100 + i, uint8=True) for i in range(2)]
    for i in range(max(1, min(args.warmup, 2))):
        obs, actions, extra = hb[i % 2]
        _, info = runner(0, state, (obs, actions, extra), with_metrics=False)
        float(info["loss"])
    sync()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        obs, actions, extra = hb[i % 2]
        _, info = runner(0, state, (obs, actions, extra), with_metrics=False)
        loss_host = float(info["loss"])  # D2H read of the step's result
    e1.record()
    sync()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / args.steps)
    h2d = model.last_h2d_bytes
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_kind = _peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    value = world * B / (ms_dev * 1e-3)
    e2e = world * B / (ms_e2e * 1e-3)
    # the dominant kernel = the tcgen05 GEMM FAMILY (one kernel template, >90 % of the step's FLOPs): every launch of the
    # profiled steps, algorithmic 2*M*N*K over the summed CUDA-event durations.  `traffic` = DRAM bytes per launch of the
    # same family from the committed ncu capture (profiles/ncu_traffic.json), averaged over one step's launches.
    epi_names = {0: "none", 1: "bias+gelu", 2: "residual", 3: "gated residual", 4: "GeGLU (dual B)", 5: "q-scale"}
    traffic, traffic_alg = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        fam = tj.get("gemm_family_per_step")
        if fam:
            traffic = fam["dram_bytes"] / fam["launches"]
            traffic_alg = fam.get("algorithmic_bytes", 0) / fam["launches"] or None
    shapes = sorted(by_shape.items(), key=lambda kv: (-kv[1][1], kv[0]))  # by time, ties by shape: deterministic
    table = [{"M": k[0], "N": k[1], "K": k[2], "batches": k[3], "epilogue": epi_names.get(k[6], k[6]),
              "launches_per_step": v[2] // nprof, "ms_per_step": v[1] / nprof,
              "tflops": v[0] / (v[1] * 1e-3) / 1e12, "frac": v[0] / (v[1] * 1e-3) / 1e12 / peak_tf} for k, v in shapes[:10]]
    roofline = {
        "bound": "tensor",
        "kernel": "gemm_bf16_tcgen05 (all launches of the step: %d per step, %.0f%% of the step time)"
                  % (round(gemm_launches), 100 * gemm_ms / ms_dev),
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "peak_kind": f"{peak_kind} sustained bf16 (kernel timed inside a long step)",
        "traffic": traffic, "algorithmic_bytes_per_launch": traffic_alg,
        "algorithmic_flops_per_launch": gemm_flops / max(gemm_launches, 1), "avg_launch_ms": gemm_ms / max(gemm_launches, 1),
        "launches_per_step": gemm_launches, "gemm_ms_per_step": gemm_ms, "gemm_share_of_step": gemm_ms / ms_dev,
        "per_shape_top10": table,
        "measured": "CUDA events around every GEMM launch on the launching stream, the same steps replayed eagerly right "
                    "after the timed region (which runs as CUDA graphs and cannot be instrumented per kernel)",
        "step_model_flops_frac": (world * B * 10.35e12 / (ms_dev * 1e-3)) / (world * peak_tf * 1e12),
    }
    infer = None
    if not args.no_infer and world == 1:  # inference is single-GPU (replicas only): reported at N = 1
        try:
            infer = infer_metrics(model, steps=100, warmup=10)
        except Exception as ex:  # pragma: no cover
            infer = {"error": repr(ex)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "lap_libero LAP-3B train step (SigLIP-So400m x2 cams + Gemma-2B + Gemma-300M expert), "
                               "per-GPU batch 32, random-init weights",
                   "per_gpu_batch": B, "global_batch": world * B, "images": "2x224x224x3", "text_tokens": 180,
                   "action_horizon": 10, "parallelism": f"dp{world}", "l2": "inputs_exceed_l2 (activations >> 126 MB)", "cuda_graph": True,
                   "e2e_images": "uint8 (loader contract), converted on the device",
                   "flops_per_sample_train": 10.35e12},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches) * args.steps,
        "gpu_launches_per_step": int(launches),
        "roofline": roofline,
        "infer": infer,
        "loss": loss_host,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference(3, 1, budget_s=60.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        except Exception as ex:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {ex}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def infer_metrics(model, *, steps: int = 100, warmup: int = 10) -> dict:
    """Second headline metric (BASELINE.json configs[3]): wall time of `sample_actions` at batch 1 — host observation
    in (uint8 camera frames, the serving wire format), host action chunk out, 10 Euler steps — p50/p90 over `steps`
    timed calls after `warmup` calls.  After warm-up the call is ONE CUDA graph: prefix pass (SigLIP +
    Gemma-2B) + the persistent denoise-loop kernel (K10).  Also the device-side split and K10's HBM roofline."""
    import torch

    from lap_b200 import ops
    from lap_b200.data import synthetic_batch
    from lap_b200.observation import Observation

    cfg = model.cfg
    b = synthetic_batch(cfg, 1, step=0, with_langact=False, uint8_images=True)
    obs = Observation.from_dict(b)
    times = []
    for i in range(warmup + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a = model.sample_actions(0, obs, num_steps=10, noise=b["noise"])
        a_host = a.cpu()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt * 1e3)
    h2d = int(model.last_h2d_bytes)
    times.sort()
    # device-side split: the whole graph, and the denoise loop alone (K10 + the V transpose), CUDA events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = model._infer_graphs.get((1, 10))
    graph_ms = None
    if g is not None:
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / 20
    denoise_ms = None
    bufs = model._bufs
    fused = model.use_denoise_megakernel and "dn.sync" in bufs
    if fused:
        x, Kc, Vc = bufs["inf.x"], bufs["inf.Kc"], bufs["inf.Vc"]
        for _ in range(3):
            model._denoise_loop_fused(x, Kc, Vc, bufs["inf.bits_s"], bufs["inf.pos_s"], 10, -0.1)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            model._denoise_loop_fused(x, Kc, Vc, bufs["inf.bits_s"], bufs["inf.pos_s"], 10, -0.1)
        e1.record(); torch.cuda.synchronize()
        denoise_ms = e0.elapsed_time(e1) / 20
    # algorithmic HBM bytes of the denoise loop: expert weights once per step, the adaRMS modulation weights once,
    # the K / V^T cache once per (step, layer)   ("each step streams the expert weights")
    e, L, nm = cfg.expert, cfg.gemma.depth, 2 * cfg.gemma.depth + 1
    per_layer = 2 * ((e.num_heads + 2) * e.head_dim * e.width + e.width * e.num_heads * e.head_dim + 3 * e.width * e.mlp_dim)
    kv = 2 * 2 * cfg.prefix_len * e.head_dim
    dn_bytes = 10 * L * (per_layer + kv) + 2 * nm * 3 * e.width * e.width
    peaks, _ = _peaks()
    return {"metric": "action-chunk infer p50 ms", "value": times[len(times) // 2], "unit": "ms",
            "p90": times[int(len(times) * 0.9)], "min": times[0], "n_gpus": 1, "steps": steps, "warmup": warmup,
            "higher_is_better": False, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "lap_libero sample_actions, batch 1, 2x224px uint8 cameras, 180-token prompt, 10 Euler steps",
                       "cuda_graph": g is not None, "fused_denoise_loop": bool(fused)},
            "device_ms": {"graph_replay": graph_ms, "denoise_loop": denoise_ms,
                          "prefix_pass": (graph_ms - denoise_ms) if (graph_ms and denoise_ms) else None},
            "e2e": {"value": times[len(times) // 2], "unit": "ms", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(a_host.numel() * 4)},
            "floor_ms": 3.4,
            "roofline": None if not denoise_ms else {
                "bound": "hbm", "kernel": "denoise_loop2_kernel (K10: 10 Euler steps x 18 expert layers, one launch)",
                "achieved": dn_bytes / (denoise_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": dn_bytes / (denoise_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": _denoise_traffic(),
                "algorithmic_bytes_per_launch": dn_bytes, "avg_launch_ms": denoise_ms}}


def _denoise_traffic():
    """DRAM bytes of one K10 launch from the committed `ncu --set full` capture (profiles/ncu_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)["denoise_loop_per_launch"]["dram_bytes"]
    except (OSError, KeyError, ValueError):
        return None


def run_infer(args) -> None:
    """`--mode infer`: the inference metric alone, as its own JSON line."""
    from lap_b200.config import get_config
    from lap_b200.model import LAP

    steps = args.steps if args.steps_given else 100
    warmup = args.warmup if args.warmup_given else 10
    model = LAP(get_config("lap_libero").model, seed=0)
    print(json.dumps(infer_metrics(model, steps=steps, warmup=warmup)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-infer", action="store_true", help="skip the inference-latency block of the train line")
    args = ap.parse_args()
    args.steps_given = any(a == "--steps" or a.startswith("--steps=") for a in sys.argv[1:])
    args.warmup_given = any(a == "--warmup" or a.startswith("--warmup=") for a in sys.argv[1:])
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "infer":
        run_infer(args)
    else:
        run_train(args)


if __name__ == "__main__":
    main()
