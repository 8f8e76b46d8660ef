"""LAP model engine: the reference's `LAP` module surface backed by liblapb200.so kernels.

Public surface (same names / argument meaning as the reference):
  LAP.compute_loss(rng, observation, actions, *, train=False, ...) -> (loss, metrics)   src/lap/models/lap.py:380-602
  LAP.sample_actions(rng, observation, *, num_steps=10, noise=None) -> actions           src/lap/models/lap.py:605-675
plus explicit `noise=` / `time=` overrides on compute_loss (jax.random's threefry stream cannot be matched).

Everything that computes runs in hand-written sm_100a kernels through the C ABI (lap_b200/ops.py); torch is the
allocator / stream owner only.  The forward keeps the activations the hand-written backward needs (no autograd, no
remat: 180 GB of HBM holds them at B=32); `forward_backward` fills a flat fp32 gradient buffer laid out like the
parameters (lap_b200/params.py).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import image_tools, ops
from . import params as P
from .config import LAPConfig
from .observation import CoTObservation, Observation, to_numpy

BF16, F32 = torch.bfloat16, torch.float32


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _bf16_round(x: float) -> float:
    return float(torch.tensor(x, dtype=torch.float32).to(torch.bfloat16).to(torch.float32))


@dataclass
class Staged:
    """Device-side inputs of one step."""

    B: int
    images: list[torch.Tensor]
    tokens: torch.Tensor  # [B, L] int32
    pm: torch.Tensor  # [B, P] uint8 prefix input mask
    par: torch.Tensor  # [B, P] uint8 prefix ar mask
    pma: torch.Tensor  # [B, P] uint8 prefix mask seen by action rows
    sm: torch.Tensor | None  # [B, A] uint8
    sar: torch.Tensor | None  # [B, A] uint8
    actions: torch.Tensor | None = None
    noise: torch.Tensor | None = None
    time: torch.Tensor | None = None
    # language-loss rows (host-computed from the host masks)
    ce_rows: torch.Tensor | None = None  # [R] int64 row index into the [B*P] prefix rows
    ce_targets: torch.Tensor | None = None  # [R] int32
    ce_weights: torch.Tensor | None = None  # [R] fp32  (gradient / loss weight of each row)
    ce_sample: torch.Tensor | None = None  # [R] int64 sample of each row (metrics)
    ce_inv_count: torch.Tensor | None = None  # [R] fp32 1/max(n_b,1) (metrics)
    R: int = 0
    n_active: float = 1.0
    n_action: float = 1.0
    sample_mask: torch.Tensor | None = None  # [B] fp32
    h2d_bytes: int = 0


class LAP:
    EOS_TOKEN = 1
    # expose for bench / tests
    last_h2d_bytes = 0

    def __init__(self, config: LAPConfig, seed: int = 0, init: bool = True, device: str | None = None,
                 reference_zero_init: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("lap_b200.LAP needs a CUDA device (sm_100a); there is no CPU fallback")
        ops.check_device()
        self.config = config
        self.cfg = config
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        cfg = config
        if not (cfg.pi05 and cfg.enable_action_training and cfg.enable_langact_training):
            raise NotImplementedError("engine implements the pi05 + action + lang-action configuration (LAP-3B)")
        if cfg.enable_vqa_training or cfg.enable_prediction_training:
            raise NotImplementedError("vqa / prediction loss heads are outside the hot path (SURVEY §8)")
        self.layout = P.FlatLayout(cfg)
        self.P = torch.zeros(self.layout.total, dtype=F32, device=self.device)
        self.W16 = torch.zeros(self.layout.total, dtype=BF16, device=self.device)
        D = cfg.gemma.width
        self.E_split = torch.zeros(cfg.vocab_size, 2 * D, dtype=BF16, device=self.device)
        self.G: torch.Tensor | None = None  # flat grads, allocated by the trainer
        # Workspaces.  `_bufs[name]` is the tensor most recently handed out under `name`; `_pool` owns EVERY tensor ever
        # handed out, keyed by (name, shape, dtype).  A workspace is therefore never freed or moved while the model
        # lives: a captured CUDA graph (training at B=32, serving at B=1, ...) keeps valid addresses even when another
        # shape is run under the same buffer names in between.
        self._bufs: dict[str, torch.Tensor] = {}
        self._pool: dict[tuple, torch.Tensor] = {}
        self._io: dict[tuple, tuple[torch.Tensor, torch.Tensor]] = {}  # persistent (pinned host, device) input buffers
        self._io_event = None
        self._resize_plans: dict = {}
        self._R_caps: dict[int, int] = {}  # batch size -> capacity of the language-loss row list (fixed shapes per B)
        self.use_cuda_graph = True
        self.use_fused_attention = True  # K1 fused tcgen05 attention forward (head_dim 256); False = GEMM+softmax+GEMM
        self.use_fused_vit_attention = os.environ.get("LAPB_FUSED_VIT", "1") != "0"  # K2 (SigLIP, head_dim 72)
        # softmax backward folded into the dP GEMM's epilogue (row term from lapb200_rowdot).  Correct (tested) but measured
        # SLOWER in the step (353.2 vs 346.4 ms, profiles/r02_softmax_bwd_fusion.md): the dP GEMM has K = head_dim (4 K
        # blocks) and is epilogue-bound already, so the extra P tile per output tile costs more than the separate
        # HBM-roofline pass it removes.  Off by default; LAPB_FUSED_SOFTMAX_BWD=1 enables it.
        self.fuse_softmax_bwd = os.environ.get("LAPB_FUSED_SOFTMAX_BWD", "0") == "1"
        self.save_mlp_act = os.environ.get("LAPB_SAVE_MLP_ACT", "1") != "0"  # keep gelu(g)*u of every layer for the backward
        # batch-1 prefix pass: deterministic split-K slabs for the down / fc2 projections (tools/gemm_small_m.py)
        self.use_small_m_split_k = os.environ.get("LAPB_SMALL_M_SPLIT_K", "1") != "0"
        self.denoise_profile = False  # accumulate per-phase ns of K10 into buf "dn.prof" (tools/denoise_prof.py)
        self.use_denoise_megakernel = True  # K10 persistent Euler-loop kernel at batch 1; False = one kernel per op
        # K10c: the 16-CTA cluster variant reads TILE-MAJOR packed copies of the expert weights and of the prefix cache
        self.denoise_cluster = os.environ.get("LAPB_DENOISE_MODE", "")[:1] == "c"
        self._packed_expert: dict | None = None
        self._infer_graphs: dict = {}
        self._infer_warm: dict = {}
        hd = cfg.gemma.head_dim
        ts = (10_000.0 ** ((2.0 / hd) * torch.arange(hd // 2, dtype=torch.float32))).to(self.device)
        self.timescale = ts
        self.ones = torch.ones(max(4096, 64), dtype=F32, device=self.device)
        self.training_saved = False
        if init:
            self.init_random(seed, reference_zero_init=reference_zero_init)

    # ------------------------------------------------------------------------------------------
    # parameters
    # ------------------------------------------------------------------------------------------
    def init_random(self, seed: int, reference_zero_init: bool = True) -> None:
        """Random init straight into the flat device buffer (same distributions as params.init_reference_params)."""
        cfg = self.cfg
        if self.layout.total < 50_000_000:
            ref = P.init_reference_params(cfg, seed, reference_zero_init=reference_zero_init)
            self.load_params(ref)
            return
        gen = torch.Generator(device=self.device).manual_seed(seed)
        lay = self.layout
        for name, shape in lay.shapes.items():
            v = lay.view(self.P, name)
            if name == "g.embed":
                v.normal_(0.0, 0.01, generator=gen)
            elif name in lay.kernel_names:
                fan_in = shape[-1]
                if name == "img.head_w" or name == "e.mod_w":
                    if reference_zero_init:
                        v.zero_()
                    else:
                        v.normal_(0.0, 0.02, generator=gen)
                elif name.startswith("img."):
                    lim = math.sqrt(6.0 / (shape[-1] + shape[-2]))
                    v.uniform_(-lim, lim, generator=gen)
                else:
                    v.normal_(0.0, 1.0 / math.sqrt(fan_in), generator=gen)
            elif name == "img.pos":
                v.normal_(0.0, 1.0 / math.sqrt(shape[-1]), generator=gen)
            elif name in ("img.ln0_s", "img.ln1_s", "img.enc_s"):
                v.fill_(1.0)
            else:
                v.zero_()
        self.refresh_compute_copy()

    def load_params(self, tree: dict) -> None:
        """Load a reference-layout tree ('/'-joined keys or nested dicts; torch or numpy leaves)."""
        flat = tree if all(isinstance(k, str) and not isinstance(v, dict) for k, v in tree.items()) else P.from_nested(tree)
        flat = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))).to(torch.float32) for k, v in flat.items()}
        want = P.reference_shapes(self.cfg)
        missing = [k for k in want if k not in flat]
        if missing:
            raise ValueError(f"parameter tree is missing {len(missing)} keys, e.g. {missing[:3]}")
        for k, s in want.items():
            if tuple(flat[k].shape) != tuple(s):
                raise ValueError(f"shape mismatch for {k}: got {tuple(flat[k].shape)}, expected {tuple(s)}")
        eng = P.reference_to_engine(self.cfg, flat)
        for name in self.layout.shapes:
            self.layout.view(self.P, name).copy_(eng[name])
        self.refresh_compute_copy()

    def params_reference(self, flat: torch.Tensor | None = None) -> dict[str, torch.Tensor]:
        """Reference-layout (CPU) view of a flat buffer (default: the master params)."""
        flat = self.P if flat is None else flat
        eng = {k: v.detach().float().cpu() for k, v in P.engine_from_flat(self.layout, flat).items()}
        return P.engine_to_reference(self.cfg, eng)

    def refresh_compute_copy(self) -> None:
        """bf16 compute copy of the master params + hi/lo split of the fp32 embedding table."""
        ops.cast_f32_bf16(self.P, self.W16)
        self.refresh_embed_split()
        self._packed_expert = None  # tile-major copies for K10c are rebuilt on their next use

    def _packed_expert_weights(self) -> dict:
        """Tile-major copies of the action expert's four per-layer weight stacks (ops.pack_tiles): what the cluster
        variant of the denoise loop streams (512 contiguous bytes per warp load).  Serving-time only; 0.6 GB."""
        if self._packed_expert is None:
            out = {}
            for name in ("e.qkv_w", "e.o_w", "e.gu_w", "e.down_w"):
                w = self.w(name)  # [L, N, K]
                L, N, K = w.shape
                dst = torch.empty_like(w)
                ops.pack_tiles(w, dst, N, K, K, 1, K, batch=L, src_bs=N * K, dst_bs=N * K)
                out[name] = dst
            self._packed_expert = out
        return self._packed_expert

    def refresh_embed_split(self) -> None:
        ops.split_hi_lo(self.p("g.embed"), self.E_split, self.cfg.vocab_size, self.cfg.gemma.width)

    def p(self, name: str, l: int | None = None) -> torch.Tensor:
        v = self.layout.view(self.P, name)
        return v if l is None else v[l]

    def w(self, name: str, l: int | None = None) -> torch.Tensor:
        v = self.layout.view(self.W16, name)
        return v if l is None else v[l]

    def g(self, name: str, l: int | None = None) -> torch.Tensor:
        v = self.layout.view(self.G, name)
        return v if l is None else v[l]

    def buf(self, name: str, shape, dtype=BF16, zero: bool = False) -> torch.Tensor:
        shape = tuple(int(s) for s in shape)
        key = (name, shape, dtype)
        t = self._pool.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            self._pool[key] = t
        self._bufs[name] = t
        return t

    def release_workspaces(self) -> None:
        """Drop every pooled workspace, staging buffer and captured inference graph (parameters stay).  Workspaces are
        never freed implicitly — a captured graph may point at them — so a long-lived process that has cycled through many
        batch shapes calls this at a quiet point; trainers drop their own graphs first (`TrainingStepRunner.reset`)."""
        torch.cuda.synchronize(self.device)
        self._infer_graphs.clear()
        self._infer_warm.clear()
        self._bufs.clear()
        self._pool.clear()
        self._io.clear()
        self._io_event = None
        self._resize_plans.clear()
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------------------------------
    # input staging (host -> device); masks are tiny and are assembled on the host
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _slab_splits(K: int, want: int) -> int:
        """Number of K-splits (<= want) for a slab split-K GEMM such that every split owns at least one 64-wide K block
        and no split is empty (the kernel requires exactly k_splits slabs); 0 if the problem is too shallow to split."""
        nk = -(-K // 64)
        for ks in range(want, 1, -1):
            per = -(-nk // ks)
            if per >= 8 and -(-nk // per) == ks:
                return ks
        return 0

    def _resize_plan(self, H: int, W: int) -> dict:
        """Device-resident filter of `resize_with_pad` for one input resolution (built once, image_tools.resize_plan)."""
        plan = self._resize_plans.get((H, W))
        if plan is None:
            S = self.cfg.image_size
            plan = image_tools.resize_plan(H, W, S, S)
            for k in ("ystart", "yw", "xstart", "xw"):
                plan[k] = torch.from_numpy(np.ascontiguousarray(plan[k])).to(self.device)
            self._resize_plans[(H, W)] = plan
        return plan

    def _stage(self, obs: Observation, actions=None, noise=None, time=None, *, with_loss: bool,
               global_counts: tuple[float, float] | None = None, aug: dict | None = None) -> Staged:
        cfg = self.cfg
        dev = self.device
        nbytes = 0
        # Inputs land in PERSISTENT device buffers (fixed addresses -> the step can be replayed as a CUDA graph) through
        # persistent pinned host buffers.  The previous step's H2D copies must have executed before the pinned
        # buffers are overwritten.
        if self._io_event is not None:
            self._io_event.synchronize()

        def up(x, dtype=None, name=None):
            nonlocal nbytes
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            dtype = dtype or t.dtype
            key = (name, tuple(t.shape), dtype)
            ent = self._io.get(key)
            if ent is None:  # keyed by shape too: a captured graph's input buffers are never reallocated
                ent = (torch.empty(tuple(t.shape), dtype=dtype).pin_memory(),
                       torch.empty(tuple(t.shape), dtype=dtype, device=dev))
                self._io[key] = ent
            host, devbuf = ent
            if t.is_cuda:
                devbuf.copy_(t)
            else:
                host.copy_(t)
                devbuf.copy_(host, non_blocking=True)
                nbytes += host.numel() * host.element_size()
            return devbuf

        images = []
        for k in cfg.image_keys:
            im = obs.images[k]
            im_t = im if isinstance(im, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(im))
            if im_t.dtype != torch.uint8:
                im_t = im_t.to(torch.float32)
            if im_t.ndim != 4 or im_t.shape[-1] != 3:
                raise ValueError(f"image {k} must be [batch, height, width, 3], got {tuple(im_t.shape)}")
            cur = up(im_t, name=f"img.{k}")
            Bi, H, W = int(im_t.shape[0]), int(im_t.shape[1]), int(im_t.shape[2])
            S = cfg.image_size
            if (H, W) != (S, S):
                # model_adapter.py:113-116: resize_with_pad, on the device (fused with the uint8 -> [-1, 1] conversion)
                plan = self._resize_plan(H, W)
                out = self.buf(f"img.rs.{k}", (Bi, S, S, 3), F32)
                ops.image_resize_pad(cur, out, Bi, H, W, S, S, plan["rh"], plan["rw"], plan["ph0"], plan["pw0"],
                                     plan["ystart"], plan["yw"], plan["ytaps"], plan["xstart"], plan["xw"], plan["xtaps"])
                cur = out
            if aug is not None and aug.get(k) is not None:
                # model_adapter.py:118-151: crop / resize / rotate / colour jitter with explicit per-sample parameters
                pa = np.zeros((Bi, 8), np.float32)
                src = np.asarray(to_numpy(aug[k]), np.float32).reshape(Bi, -1)
                pa[:, : src.shape[1]] = src
                out = self.buf(f"img.aug.{k}", (Bi, S, S, 3), F32)
                ops.image_augment(cur, out, Bi, S, S, up(pa, name=f"aug.{k}"))
                cur = out
            images.append(cur)
        B = images[0].shape[0]
        Np, L, C = cfg.num_patches, cfg.max_token_len, len(cfg.image_keys)
        tokens = to_numpy(obs.tokenized_prompt).astype(np.int32)
        prompt_mask = to_numpy(obs.tokenized_prompt_mask).astype(bool)
        la = getattr(obs, "tokenized_langact_mask", None)
        la = None if la is None else to_numpy(la).astype(bool)
        img_masks = []
        for k in cfg.image_keys:
            m = obs.image_masks.get(k) if obs.image_masks is not None else None
            m = np.ones(B, dtype=bool) if m is None else to_numpy(m).astype(bool)
            img_masks.append(np.repeat(m[:, None], Np, axis=1))
        pm = np.concatenate(img_masks + [prompt_mask], axis=1)  # lap.py:134-151
        par = np.concatenate([np.zeros((B, C * Np), bool), la if la is not None else np.zeros((B, L), bool)], axis=1)
        if la is not None:  # lap.py:303-325
            pma = pm & ~np.concatenate([np.zeros((B, C * Np), bool), la], axis=1)
        else:
            pma = pm
        A = cfg.action_horizon
        st = Staged(B=B, images=images, tokens=up(tokens, name="tokens"), pm=up(pm.astype(np.uint8), name="pm"),
                    par=up(par.astype(np.uint8), name="par"), pma=up(pma.astype(np.uint8), name="pma"),
                    sm=up(np.ones((B, A), np.uint8), name="sm"),
                    sar=up(np.tile(np.array([1] + [0] * (A - 1), np.uint8), (B, 1)), name="sar"))
        if actions is not None:
            st.actions = up(to_numpy(actions).astype(np.float32), name="actions")
        if noise is not None:
            st.noise = up(to_numpy(noise).astype(np.float32), name="noise")
        if time is not None:
            st.time = up(to_numpy(time).astype(np.float32), name="time")
        if with_loss:
            if la is None:
                raise ValueError("compute_loss needs tokenized_langact_mask (lap.py:232)")
            loss_mask_in = getattr(obs, "token_loss_mask", None)
            tlm = np.ones((B, L), bool) if loss_mask_in is None else to_numpy(loss_mask_in).astype(bool)
            sm_in = getattr(obs, "sample_mask", None)
            smask = np.ones(B, bool) if sm_in is None else to_numpy(sm_in).astype(bool)
            lm = la[:, 1:] & prompt_mask[:, 1:] & tlm[:, 1:] & smask[:, None]  # lap.py:231-237
            n_b = np.maximum(lm.sum(-1), 1).astype(np.float32)
            n_active_local = float(smask.sum()) if sm_in is not None else float(B)
            n_active, n_action = global_counts if global_counts is not None else (n_active_local, float(B))
            n_active = max(n_active, 1.0)
            bb, jj = np.nonzero(lm)
            R = len(bb)
            # fixed row capacity (padding rows weigh 0) so that the captured step keeps its shapes; grows if needed
            if B not in self._R_caps or R > self._R_caps[B]:
                self._R_caps[B] = max(_round_up(int(R * 1.25) + 1, 128), 128)
            Rp = self._R_caps[B]
            rows = np.zeros(Rp, np.int64)
            tgt = np.zeros(Rp, np.int32)
            wts = np.zeros(Rp, np.float32)
            smp = np.zeros(Rp, np.int64)
            inv = np.zeros(Rp, np.float32)
            rows[:R] = bb * cfg.prefix_len + C * Np + jj  # pre_logits[:, :-1][:, -(L-1):] -> prefix position C*Np + j
            tgt[:R] = tokens[bb, jj + 1]
            wts[:R] = cfg.language_loss_weight / (n_b[bb] * n_active)
            smp[:R] = bb
            inv[:R] = 1.0 / n_b[bb]
            st.ce_rows, st.ce_targets, st.ce_weights = up(rows, name="ce_rows"), up(tgt, name="ce_tgt"), up(wts, name="ce_w")
            st.ce_sample, st.ce_inv_count = up(smp, name="ce_smp"), up(inv, name="ce_inv")
            st.R = Rp
            st.n_active, st.n_action = n_active, max(n_action, 1.0)
            st.sample_mask = up(smask.astype(np.float32), name="sample_mask")
        self._io_event = torch.cuda.Event()
        self._io_event.record()
        st.h2d_bytes = nbytes
        LAP.last_h2d_bytes = nbytes
        return st

    # ------------------------------------------------------------------------------------------
    # GEMM helpers.  Weights are [out, in] (K-major B of the forward GEMM).
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _dgrad(dY, W, dX, M, N, K, **kw):
        """dX[M,K] = dY[M,N] @ W[N,K]  (B operand read MN-major)."""
        ops.gemm(dY, W, dX, M=M, N=K, K=N, b_major=1, lda=N, ldb=K, ldc=K, **kw)

    @staticmethod
    def _wgrad(dY, X, dW, M, N, K, ldx=None, lddy=None, accumulate=False):
        """dW[N,K] = dY[M,N]^T @ X[M,K]  (both operands read MN-major, fp32 out)."""
        ops.gemm(dY, X, dW, M=N, N=K, K=M, a_major=1, b_major=1, lda=lddy if lddy is not None else N,
                 ldb=ldx if ldx is not None else K, ldc=K, accumulate=accumulate)

    @staticmethod
    def _lin(A, W, C, M, N, K, **kw):
        """C = epi(A @ W^T).  M <= 16 rows (one denoise step at batch 1) take the weight-streaming kernel."""
        if M <= 16 and K % 32 == 0 and kw.get("epi", ops.EPI_NONE) in (ops.EPI_NONE, ops.EPI_RESID,
                                                                       ops.EPI_GATED_RESID, ops.EPI_GEGLU):
            ops.skinny_gemm(A, W, C, M=M, N=N, K=K, epi=kw.get("epi", ops.EPI_NONE), bias=kw.get("bias"),
                            resid=kw.get("resid"), gate=kw.get("gate"), ldg=kw.get("ldg", 0),
                            gate_rows=kw.get("gate_rows", 1), Y2=kw.get("C2"), ldy2=kw.get("ldc2", 0))
        else:
            ops.gemm(A, W, C, M=M, N=N, K=K, **kw)

    # ------------------------------------------------------------------------------------------
    # SigLIP tower (OP/models/siglip.py), all cameras in one pass, rows ordered (b, cam, patch)
    # ------------------------------------------------------------------------------------------
    def _siglip_fwd(self, st: Staged, X0: torch.Tensor, rows_per_sample: int, softmax_mode: int = 0,
                    save: bool = True) -> None:
        """`save=False` (inference): the attention probabilities are not written out."""
        cfg, s = self.cfg, self.cfg.siglip
        B, C, Np, W, F, nh, hd = st.B, len(cfg.image_keys), cfg.num_patches, s.width, s.mlp_dim, s.num_heads, s.head_dim
        D = cfg.gemma.width
        Ni, Ms = B * C, B * C * Np
        pk = s.patch_size * s.patch_size * 3
        patches = self.buf("img.patches", (Ms, pk), F32)
        pk_pad = _round_up(pk, 8)
        p_hi = self.buf("img.patches_hi", (Ms, pk_pad), zero=True)
        p_lo = self.buf("img.patches_lo", (Ms, pk_pad), zero=True)
        ops.patchify(st.images, patches, B, C, cfg.image_size, cfg.image_size, s.patch_size, p_hi, p_lo, pk_pad)
        x = self.buf("img.x.0", (Ms, W))
        ops.sgemm(patches, self.p("img.patch_w"), x, Ms, W, pk, pk, 1, pk, 1, ldc=W, bias=self.p("img.patch_b"),
                  table=self.p("img.pos"), table_rows=Np)
        q_div = _bf16_round(math.sqrt(hd))
        # batch-1 inference: fc2 (M = 512 rows, K = 4304) has 36 output tiles for 148 SMs -> deterministic split-K slabs,
        # summed (+ bias + residual, same rounding points) inside the NEXT LayerNorm kernel (ops.resid_norm_fwd)
        ks2 = self._slab_splits(F, 4) if (not save and Ms <= 1024 and self.use_small_m_split_k) else 0
        acc2 = self.buf("img.acc2", (max(ks2, 1), Ms, W), F32) if ks2 else None
        pending = None  # (residual x1, bias) of an fc2 whose slabs still wait in acc2
        for l in range(s.depth):
            y0 = self.buf(f"img.y0.{l}", (Ms, W))
            mean0, rstd0 = self.buf(f"img.mean0.{l}", (Ms,), F32), self.buf(f"img.rstd0.{l}", (Ms,), F32)
            if pending is not None:
                x = self.buf(f"img.x.{l}", (Ms, W))
                ops.resid_norm_fwd(pending[0], acc2, ks2, Ms * W, pending[1], x, True, self.p("img.ln0_s", l),
                                   self.p("img.ln0_b", l), y0, mean0, rstd0, Ms, W)
                pending = None
            else:
                ops.layernorm_fwd(x, self.p("img.ln0_s", l), self.p("img.ln0_b", l), y0, mean0, rstd0, Ms, W)
            qkv = self.buf(f"img.qkv.{l}", (Ms, 3 * W))
            ops.gemm(y0, self.w("img.qkv_w", l), qkv, M=Ms, N=3 * W, K=W, bias=self.p("img.qkv_b", l),
                     epi=ops.EPI_QSCALE, q_cols=W, q_div=q_div)
            o = self.buf(f"img.o.{l}", (Ms, W))
            if self.use_fused_vit_attention and 64 < hd <= 80:
                # K2: S = QK^T -> softmax -> PV in one tcgen05 kernel; P is written out only for the backward pass
                Pm = self.buf(f"img.P.{l}", (Ni, nh, Np, Np)) if save else None
                ops.vit_attn_fwd(qkv, o, Pm, Ni, nh, Np, hd, softmax_mode)
            else:
                Pm = self.buf(f"img.P.{l}", (Ni, nh, Np, Np))
                qf = qkv.view(-1)
                ops.gemm(qf, qf[W:], Pm, M=Np, N=Np, K=hd, lda=3 * W, ldb=3 * W, ldc=Np, batch_i=nh, batch_o=Ni,
                         a_bs=(hd, Np * 3 * W), b_bs=(hd, Np * 3 * W), c_bs=(Np * Np, nh * Np * Np))
                ops.vit_softmax_fwd(Pm, Ni * nh * Np, Np, Np, softmax_mode)
                ops.gemm(Pm, qf[2 * W:], o, M=Np, N=hd, K=Np, b_major=1, lda=Np, ldb=3 * W, ldc=W, batch_i=nh,
                         batch_o=Ni, a_bs=(Np * Np, nh * Np * Np), b_bs=(hd, Np * 3 * W), c_bs=(hd, Np * W))
            x1 = self.buf(f"img.x1.{l}", (Ms, W))
            ops.gemm(o, self.w("img.out_w", l), x1, M=Ms, N=W, K=W, bias=self.p("img.out_b", l), epi=ops.EPI_RESID,
                     resid=x)
            y1 = self.buf(f"img.y1.{l}", (Ms, W))
            mean1, rstd1 = self.buf(f"img.mean1.{l}", (Ms,), F32), self.buf(f"img.rstd1.{l}", (Ms,), F32)
            ops.layernorm_fwd(x1, self.p("img.ln1_s", l), self.p("img.ln1_b", l), y1, mean1, rstd1, Ms, W)
            hpre, hact = self.buf(f"img.hpre.{l}", (Ms, F)), self.buf(f"img.hact.{l}", (Ms, F))
            ops.gemm(y1, self.w("img.fc1_w", l), hact, M=Ms, N=F, K=W, bias=self.p("img.fc1_b", l),
                     epi=ops.EPI_BIAS_GELU, C2=hpre, ldc2=F)
            if ks2:
                ops.gemm(hact, self.w("img.fc2_w", l), acc2, M=Ms, N=W, K=F, k_splits=ks2, split_stride=Ms * W,
                         cta_group=1, block_n=128)
                pending = (x1, self.p("img.fc2_b", l))
            else:
                x2 = self.buf(f"img.x.{l + 1}", (Ms, W))
                ops.gemm(hact, self.w("img.fc2_w", l), x2, M=Ms, N=W, K=F, bias=self.p("img.fc2_b", l), epi=ops.EPI_RESID,
                         resid=x1)
                x = x2
        yenc = self.buf("img.yenc", (Ms, W))
        if pending is not None:
            ops.resid_norm_fwd(pending[0], acc2, ks2, Ms * W, pending[1], self.buf(f"img.x.{s.depth}", (Ms, W)), True,
                               self.p("img.enc_s"), self.p("img.enc_b"), yenc, self.buf("img.mean_e", (Ms,), F32),
                               self.buf("img.rstd_e", (Ms,), F32), Ms, W)
        else:
            ops.layernorm_fwd(x, self.p("img.enc_s"), self.p("img.enc_b"), yenc, self.buf("img.mean_e", (Ms,), F32),
                              self.buf("img.rstd_e", (Ms,), F32), Ms, W)
        # head Dense -> image tokens written straight into the prefix rows [b, cam*Np + t]
        ops.gemm(yenc, self.w("img.head_w"), X0, M=Np, N=D, K=W, bias=self.p("img.head_b"), batch_i=C, batch_o=B,
                 a_bs=(Np * W, C * Np * W), c_bs=(Np * D, rows_per_sample * D), ldc=D)

    def _siglip_bwd(self, st: Staged, dX0: torch.Tensor, rows_per_sample: int) -> None:
        self._siglip_bwd_head(st, dX0, rows_per_sample)
        self._siglip_bwd_layers(st, self.cfg.siglip.depth, 0)
        self._siglip_bwd_tail(st)

    def _siglip_bwd_head(self, st: Staged, dX0: torch.Tensor, rows_per_sample: int) -> None:
        """head Dense + encoder_norm backward; leaves d(last block output) in `img.dx`."""
        cfg, s = self.cfg, self.cfg.siglip
        B, C, Np, W, F, nh, hd = st.B, len(cfg.image_keys), cfg.num_patches, s.width, s.mlp_dim, s.num_heads, s.head_dim
        D = cfg.gemma.width
        Ni, Ms = B * C, B * C * Np
        pk = s.patch_size * s.patch_size * 3
        q_div = _bf16_round(math.sqrt(hd))
        # image rows of dX0 -> contiguous [Ms, D]
        dximg = self.buf("img.dximg", (Ms, D))
        dximg.view(B, C * Np * D).copy_(dX0.view(B, rows_per_sample * D)[:, : C * Np * D])
        yenc = self._bufs["img.yenc"]
        self._wgrad(dximg, yenc, self.g("img.head_w"), Ms, D, W)
        ops.colsum(dximg, D, self.g("img.head_b"), Ms, D)
        dy = self.buf("img.dy", (Ms, W))
        self._dgrad(dximg, self.w("img.head_w"), dy, Ms, D, W)
        dx = self.buf("img.dx", (Ms, W))
        ops.layernorm_bwd(dy, self._bufs[f"img.x.{s.depth}"], self.p("img.enc_s"), self._bufs["img.mean_e"],
                          self._bufs["img.rstd_e"], None, dx, self.g("img.enc_s"), self.g("img.enc_b"), Ms, W)

    def _siglip_bwd_layers(self, st: Staged, l_hi: int, l_lo: int) -> None:
        """Encoder blocks l_hi-1 ... l_lo; their weight gradients are final on return."""
        cfg, s = self.cfg, self.cfg.siglip
        B, C, Np, W, F, nh, hd = st.B, len(cfg.image_keys), cfg.num_patches, s.width, s.mlp_dim, s.num_heads, s.head_dim
        D = cfg.gemma.width
        Ni, Ms = B * C, B * C * Np
        pk = s.patch_size * s.patch_size * 3
        q_div = _bf16_round(math.sqrt(hd))
        dy = self.buf("img.dy", (Ms, W))
        dx = self.buf("img.dx", (Ms, W))
        dh = self.buf("img.dh", (Ms, F))
        dqkv = self.buf("img.dqkv", (Ms, 3 * W))
        do = self.buf("img.do", (Ms, W))
        dP = self.buf("img.dP", (Ni, nh, Np, Np))
        for l in reversed(range(l_lo, l_hi)):
            g = lambda n: self.g(n, l)
            hact, hpre, y1, x1 = (self._bufs[f"img.{n}.{l}"] for n in ("hact", "hpre", "y1", "x1"))
            # fc2
            self._wgrad(dx, hact, g("img.fc2_w"), Ms, W, F)
            ops.colsum(dx, W, g("img.fc2_b"), Ms, W)
            self._dgrad(dx, self.w("img.fc2_w", l), dh, Ms, W, F)
            ops.gelu_bwd(dh, hpre, Ms * F)
            # fc1
            self._wgrad(dh, y1, g("img.fc1_w"), Ms, F, W)
            ops.colsum(dh, F, g("img.fc1_b"), Ms, F)
            self._dgrad(dh, self.w("img.fc1_w", l), dy, Ms, F, W)
            ops.layernorm_bwd(dy, x1, self.p("img.ln1_s", l), self._bufs[f"img.mean1.{l}"], self._bufs[f"img.rstd1.{l}"],
                              dx, dx, g("img.ln1_s"), g("img.ln1_b"), Ms, W)
            # out projection
            o, qkv, Pm = (self._bufs[f"img.{n}.{l}"] for n in ("o", "qkv", "P"))
            self._wgrad(dx, o, g("img.out_w"), Ms, W, W)
            ops.colsum(dx, W, g("img.out_b"), Ms, W)
            self._dgrad(dx, self.w("img.out_w", l), do, Ms, W, W)
            # attention backward, batched over (image, head)
            qf, dqf, dof = qkv.view(-1), dqkv.view(-1), do.view(-1)
            bs_qkv, bs_o, bs_p = (hd, Np * 3 * W), (hd, Np * W), (Np * Np, nh * Np * Np)
            # dP = dO V^T
            if self.fuse_softmax_bwd:  # dS straight from the dP GEMM's epilogue (see the Gemma backward)
                delta = self.buf("img.delta", (Ni, nh, Np), F32)
                ops.rowdot(do, o, delta, Np, hd, W, W, nbi=nh, nbo=Ni, d_bs=(hd, Np * W), o_bs=(hd, Np * W))
                ops.gemm(dof, qf[2 * W:], dP, M=Np, N=Np, K=hd, lda=W, ldb=3 * W, ldc=Np, batch_i=nh, batch_o=Ni,
                         a_bs=bs_o, b_bs=bs_qkv, c_bs=bs_p, epi=ops.EPI_SOFTMAX_BWD, C2=Pm, ldc2=Np, bias=delta)
            else:
                ops.gemm(dof, qf[2 * W:], dP, M=Np, N=Np, K=hd, lda=W, ldb=3 * W, ldc=Np, batch_i=nh, batch_o=Ni,
                         a_bs=bs_o, b_bs=bs_qkv, c_bs=bs_p)
            # dV = P^T dO
            ops.gemm(Pm, dof, dqf[2 * W:], M=Np, N=hd, K=Np, a_major=1, b_major=1, lda=Np, ldb=W, ldc=3 * W,
                     batch_i=nh, batch_o=Ni, a_bs=bs_p, b_bs=bs_o, c_bs=bs_qkv)
            if not self.fuse_softmax_bwd:
                ops.softmax_bwd(Pm, dP, dP, Ni * nh * Np, Np)
            # dQ_pre = (dS K) / q_div
            ops.gemm(dP, qf[W:], dqf, M=Np, N=hd, K=Np, b_major=1, lda=Np, ldb=3 * W, ldc=3 * W, batch_i=nh, batch_o=Ni,
                     a_bs=bs_p, b_bs=bs_qkv, c_bs=bs_qkv, epi=ops.EPI_QSCALE, q_cols=hd, q_div=q_div)
            # dK = dS^T Q_scaled
            ops.gemm(dP, qf, dqf[W:], M=Np, N=hd, K=Np, a_major=1, b_major=1, lda=Np, ldb=3 * W, ldc=3 * W,
                     batch_i=nh, batch_o=Ni, a_bs=bs_p, b_bs=bs_qkv, c_bs=bs_qkv)
            y0, x = self._bufs[f"img.y0.{l}"], self._bufs[f"img.x.{l}"]
            self._wgrad(dqkv, y0, g("img.qkv_w"), Ms, 3 * W, W)
            ops.colsum(dqkv, 3 * W, g("img.qkv_b"), Ms, 3 * W)
            self._dgrad(dqkv, self.w("img.qkv_w", l), dy, Ms, 3 * W, W)
            ops.layernorm_bwd(dy, x, self.p("img.ln0_s", l), self._bufs[f"img.mean0.{l}"], self._bufs[f"img.rstd0.{l}"],
                              dx, dx, g("img.ln0_s"), g("img.ln0_b"), Ms, W)

    def _siglip_bwd_tail(self, st: Staged) -> None:
        """Patch-embedding gradients (position table, bias, fp32 conv kernel)."""
        cfg, s = self.cfg, self.cfg.siglip
        B, C, Np, W, F, nh, hd = st.B, len(cfg.image_keys), cfg.num_patches, s.width, s.mlp_dim, s.num_heads, s.head_dim
        D = cfg.gemma.width
        Ni, Ms = B * C, B * C * Np
        pk = s.patch_size * s.patch_size * 3
        q_div = _bf16_round(math.sqrt(hd))
        dx = self.buf("img.dx", (Ms, W))
        # patch embedding: pos (sum over images), bias, kernel (fp32)
        ops.colsum(dx, Np * W, self.g("img.pos"), Ni, Np * W)
        ops.colsum(dx, W, self.g("img.patch_b"), Ms, W)
        # fp32 conv kernel gradient on the tensor cores: dW = dx^T [patch_hi + patch_lo] (bf16 hi/lo split of the fp32
        # patch matrix keeps ~16 mantissa bits; siglip.py:216-223 evaluates this conv in fp32)
        pk_pad = _round_up(pk, 8)
        dwp = self.buf("img.dpatch_w", (W, pk_pad), F32)
        ops.gemm(dx, self._bufs["img.patches_hi"], dwp, M=W, N=pk_pad, K=Ms, a_major=1, b_major=1, lda=W, ldb=pk_pad,
                 ldc=pk_pad)
        ops.gemm(dx, self._bufs["img.patches_lo"], dwp, M=W, N=pk_pad, K=Ms, a_major=1, b_major=1, lda=W, ldb=pk_pad,
                 ldc=pk_pad, accumulate=True)
        self.g("img.patch_w").copy_(dwp[:, :pk])

    # ------------------------------------------------------------------------------------------
    # suffix (flow-matching) embedding — fp32 (OP/models/pi0.py:139-186, lap.py:185-207)
    # ------------------------------------------------------------------------------------------
    def _suffix_embed(self, st: Staged, x_t: torch.Tensor | None, time: torch.Tensor, XE: torch.Tensor) -> None:
        """action_in_proj(x_t) -> XE (bf16); time MLP -> cond (fp32) and cond16 (bf16).  If x_t is None it is
        built from (actions, noise, time) together with u_t."""
        cfg = self.cfg
        B, A, ad, D1 = st.B, cfg.action_horizon, cfg.action_dim, cfg.expert.width
        te = self.buf("suf.te", (B, D1), F32)
        if x_t is None:
            x_t = self.buf("suf.x_t", (B, A * ad), F32)
            u_t = self.buf("suf.u_t", (B, A * ad), F32)
            ops.suffix_inputs(st.actions, st.noise, time, x_t, u_t, te, B, A * ad, D1)
        else:
            ops.suffix_inputs(None, None, time, None, None, te, B, A * ad, D1)
        ops.linear_f32(x_t, self.p("action_in_w"), XE, B * A, D1, ad, bias=self.p("action_in_b"))
        z1, s1 = self.buf("suf.z1", (B, D1), F32), self.buf("suf.s1", (B, D1), F32)
        z2, cond = self.buf("suf.z2", (B, D1), F32), self.buf("suf.cond", (B, D1), F32)
        cond16 = self.buf("suf.cond16", (B, D1))
        ops.linear_f32(te, self.p("time_in_w"), z1, B, D1, D1, bias=self.p("time_in_b"))
        ops.swish_fwd(z1, s1, None, B * D1)
        ops.linear_f32(s1, self.p("time_out_w"), z2, B, D1, D1, bias=self.p("time_out_b"))
        ops.swish_fwd(z2, cond, cond16, B * D1)
        # all 2L+1 adaRMS modulation Dense layers in one GEMM: mod[b, i*3D1 : (i+1)*3D1]
        nm = P.n_mod(cfg)
        mod = self.buf("suf.mod", (B, nm * 3 * D1))
        self._lin(cond16, self.w("e.mod_w"), mod, B, nm * 3 * D1, D1, bias=self.p("e.mod_b").view(-1))

    # ------------------------------------------------------------------------------------------
    # Gemma multi-expert transformer (src/lap/models/backbones/gemma.py:455-531)
    # ------------------------------------------------------------------------------------------
    def _gemma_fwd(self, B: int, X, XE, bits, positions, *, save: bool, kv_cache=None, write_cache=None,
                   tag: str = "g"):
        """Runs the depth-L stack.  X: prefix rows [B*P, D] or None; XE: suffix rows [B*A, D1] or None.
        kv_cache=(K,V) [L,B,Tpad,hd] already holding the prefix (suffix-only pass); write_cache=(K,V) to fill
        (prefix-only pass).  Returns (X_out, XE_out) — the residual streams BEFORE the final norm."""
        cfg, g, e = self.cfg, self.cfg.gemma, self.cfg.expert
        Pn = cfg.prefix_len if (X is not None or kv_cache is not None) else 0
        A = cfg.action_horizon if XE is not None else 0
        T = Pn + A
        Tpad = _round_up(cfg.prefix_len + cfg.action_horizon, 64)
        W32 = Tpad // 32
        D, D1, hd, NH = g.width, e.width, g.head_dim, g.num_heads
        QKV = (NH + 2) * hd
        F, F1 = g.mlp_dim, e.mlp_dim
        Mg, Me = B * Pn if X is not None else 0, B * A
        t_begin = Pn if X is None else 0
        Tq = T - t_begin
        nm3 = P.n_mod(cfg) * 3 * D1
        mod = self._bufs.get("suf.mod")
        sv = (lambda n, l: f"{tag}.{n}.{l}") if save else (lambda n, l: f"{tag}.{n}.tmp")
        qscale = hd ** -0.5
        for l in range(g.depth):
            if kv_cache is not None:
                Kc, Vc = kv_cache[0][l], kv_cache[1][l]
            elif write_cache is not None:
                Kc, Vc = write_cache[0][l], write_cache[1][l]
            else:
                Kc = self.buf(sv("Kc", l), (B, Tpad, hd), zero=True)
                Vc = self.buf(sv("Vc", l), (B, Tpad, hd), zero=True)
            qkv0 = qkv1 = None
            if X is not None:
                h = self.buf(sv("h", l), (Mg, D))
                rstd = self.buf(sv("rstd", l), (Mg,), F32)
                ops.rmsnorm_fwd(X, h, rstd, Mg, D, scale=self.p("g.attn_norm_s", l))
                qkv0 = self.buf(f"{tag}.qkv0", (Mg, QKV))
                ops.gemm(h, self.w("g.qkv_w", l), qkv0, M=Mg, N=QKV, K=D)
            if XE is not None:
                hE = self.buf(sv("hE", l), (Me, D1))
                rstdE = self.buf(sv("rstdE", l), (Me,), F32)
                ops.rmsnorm_fwd(XE, hE, rstdE, Me, D1, mod=mod.view(-1)[(2 * l) * 3 * D1:], ldmod=nm3, rows_per_sample=A)
                qkv1 = self.buf(f"{tag}.qkv1", (Me, QKV))
                self._lin(hE, self.w("e.qkv_w", l), qkv1, Me, QKV, D1)
            Q = self.buf(sv("Q", l), (B, Tq, NH, hd))
            ops.rope_fwd(qkv0, qkv1, positions, self.timescale, Q, Kc, Vc, B, Pn, A, Tpad, NH, hd, t_begin, qscale)
            R = Tq * NH
            fused = (hd == 256 and self.use_fused_attention and not (X is None and kv_cache is not None and Me <= 16))
            if fused:
                # K1: fused tcgen05 attention; writes O straight into the per-expert buffers (and P for the backward)
                O0 = self.buf(sv("O0", l), (Mg, NH * hd)) if X is not None else None
                O1f = self.buf(sv("O1", l), (Me, NH * hd)) if XE is not None else None
                Pm = self.buf(sv("P", l), (B, R, Tpad)) if save else None
                split = (Pn * NH) if X is not None else 0
                ops.fa_gemma_fwd(Q, Kc, Vc, bits, Pm, O0 if O0 is not None else O1f, O1f if O1f is not None else O0,
                                 B, R, NH, Tq, T, Tpad, W32, split, hd)
                S = None
                Oc = None
            elif X is None and kv_cache is not None and Me <= 16:
                # suffix-only denoise step: a handful of query tokens against the cache
                Oc = self.buf(f"{tag}.Oc", (B, R, hd))
                ops.decode_attn(Q, Kc, Vc, bits, Oc, B, Tq, NH, hd, T, Tpad, W32)
                S = None
            else:
                S = self.buf(f"{tag}.S", (B, R, Tpad), F32)
            if S is not None:
                ops.gemm(Q, Kc, S, M=R, N=Tpad, K=hd, ldc=Tpad, batch_i=B, a_bs=(R * hd, 0), b_bs=(Tpad * hd, 0),
                         c_bs=(R * Tpad, 0))
                Pm = self.buf(sv("P", l), (B, R, Tpad))
                ops.attn_softmax_fwd(S, bits, Pm, B, R, NH, T, Tpad, W32)
                Oc = self.buf(f"{tag}.Oc", (B, R, hd))
                ops.gemm(Pm, Vc, Oc, M=R, N=hd, K=Tpad, b_major=1, lda=Tpad, ldb=hd, ldc=hd, batch_i=B,
                         a_bs=(R * Tpad, 0), b_bs=(Tpad * hd, 0), c_bs=(R * hd, 0))
            if X is not None:
                if not fused:
                    O0 = self.buf(sv("O0", l), (Mg, NH * hd))
                    O0.view(B, Pn * NH * hd).copy_(Oc.view(B, R * hd)[:, : Pn * NH * hd])
                X1 = self.buf(sv("X1", l), (Mg, D))
                ops.gemm(O0, self.w("g.o_w", l), X1, M=Mg, N=D, K=NH * hd, epi=ops.EPI_RESID, resid=X)
                h2 = self.buf(sv("h2", l), (Mg, D))
                rstd2 = self.buf(sv("rstd2", l), (Mg,), F32)
                ops.rmsnorm_fwd(X1, h2, rstd2, Mg, D, scale=self.p("g.ffn_norm_s", l))
                # `act` is kept per layer when the backward follows: geglu_bwd then need not rewrite it (one sixth of its
                # HBM traffic), at 0.7 GB per layer of the 180 GB
                act = self.buf(sv("act", l) if (save and self.save_mlp_act) else f"{tag}.act", (Mg, F))
                GU = self.buf(sv("GU", l), (Mg, 2 * F))
                ops.gemm(h2, self.w("g.gu_w", l), act, M=Mg, N=F, K=D, epi=ops.EPI_GEGLU, C2=GU, ldc2=2 * F)
                X2 = self.buf(sv("X", l + 1), (Mg, D))
                ops.gemm(act, self.w("g.down_w", l), X2, M=Mg, N=D, K=F, epi=ops.EPI_RESID, resid=X1)
                X = X2
            if XE is not None:
                if fused:
                    O1 = O1f
                elif X is not None or Pn == 0 or t_begin == 0:
                    O1 = self.buf(sv("O1", l), (Me, NH * hd))
                    O1.view(B, A * NH * hd).copy_(Oc.view(B, R * hd)[:, (Tq - A) * NH * hd:])
                else:
                    O1 = Oc.view(Me, NH * hd)
                gate_a = mod.view(-1)[(2 * l) * 3 * D1 + 2 * D1:]
                gate_f = mod.view(-1)[(2 * l + 1) * 3 * D1 + 2 * D1:]
                XE1 = self.buf(sv("XE1", l), (Me, D1))
                yEa = self.buf(sv("yEa", l), (Me, D1))
                self._lin(O1, self.w("e.o_w", l), XE1, Me, D1, NH * hd, epi=ops.EPI_GATED_RESID, resid=XE,
                          gate=gate_a, ldg=nm3, gate_rows=A, C2=yEa if save else None, ldc2=D1)
                hE2 = self.buf(sv("hE2", l), (Me, D1))
                rstdE2 = self.buf(sv("rstdE2", l), (Me,), F32)
                ops.rmsnorm_fwd(XE1, hE2, rstdE2, Me, D1, mod=mod.view(-1)[(2 * l + 1) * 3 * D1:], ldmod=nm3,
                                rows_per_sample=A)
                actE = self.buf(sv("actE", l), (Me, F1))
                GUE = self.buf(sv("GUE", l), (Me, 2 * F1))
                self._lin(hE2, self.w("e.gu_w", l), actE, Me, F1, D1, epi=ops.EPI_GEGLU, C2=GUE if save else None,
                          ldc2=2 * F1)
                XE2 = self.buf(sv("XE", l + 1), (Me, D1))
                yEf = self.buf(sv("yEf", l), (Me, D1))
                self._lin(actE, self.w("e.down_w", l), XE2, Me, D1, F1, epi=ops.EPI_GATED_RESID, resid=XE1,
                          gate=gate_f, ldg=nm3, gate_rows=A, C2=yEf if save else None, ldc2=D1)
                XE = XE2
        return X, XE

    # ------------------------------------------------------------------------------------------
    # training forward (+ loss)
    # ------------------------------------------------------------------------------------------
    def _forward_loss(self, st: Staged, *, save: bool, compute_grad_seed: bool, softmax_mode: int = 0):
        cfg, g, e = self.cfg, self.cfg.gemma, self.cfg.expert
        B, Pn, A, L = st.B, cfg.prefix_len, cfg.action_horizon, cfg.max_token_len
        C, Np = len(cfg.image_keys), cfg.num_patches
        D, D1, ad, V = g.width, e.width, cfg.action_dim, cfg.vocab_size
        T = Pn + A
        Tpad = _round_up(T, 64)
        Mg, Me = B * Pn, B * A
        sv0 = "g.X.0" if save else "g.X.tmp0"
        X0 = self.buf(sv0, (Mg, D))
        self._siglip_fwd(st, X0, Pn, softmax_mode, save=save)
        ops.embed_fwd(st.tokens, self.p("g.embed"), X0, B, L, C * Np, Pn, D, math.sqrt(D))
        XE0 = self.buf("g.XE.0" if save else "g.XE.tmp0", (Me, D1))
        self._suffix_embed(st, None, st.time, XE0)
        bits = self.buf("mask.bits", (B, T, Tpad // 32), torch.int32)
        positions = self.buf("mask.pos", (B, T), torch.int32)
        ops.mask_build(st.pm, st.par, st.pma, st.sm, st.sar, bits, positions, B, Pn, A, Tpad // 32)
        X, XE = self._gemma_fwd(B, X0, XE0, bits, positions, save=save)
        # ---- language loss: only rows that carry loss go through final norm + LM head (masked rows weigh 0) ----
        R = st.R
        pre2 = self.buf("loss.pre2", (R, 2 * D))
        rstdF = self.buf("loss.rstdF", (R,), F32)
        ops.rmsnorm_fwd(X, pre2, rstdF, R, D, scale=self.p("g.final_norm_s"), row_idx=st.ce_rows, ldy=2 * D, dup=True)
        logits = self.buf("loss.logits", (R, V), F32)
        ops.gemm(pre2, self.E_split, logits, M=R, N=V, K=2 * D)
        nll = self.buf("loss.nll", (R,), F32)
        dlogits = self.buf("loss.dlogits", (R, V)) if compute_grad_seed else None
        ops.ce_fwd_bwd(logits, V, st.ce_targets, st.ce_weights, nll, dlogits, V, R, V)
        # ---- action loss ----
        nm3 = P.n_mod(cfg) * 3 * D1
        mod = self._bufs["suf.mod"]
        sufout = self.buf("loss.sufout", (Me, D1))
        rstdEF = self.buf("loss.rstdEF", (Me,), F32)
        ops.rmsnorm_fwd(XE, sufout, rstdEF, Me, D1, mod=mod.view(-1)[(P.n_mod(cfg) - 1) * 3 * D1:], ldmod=nm3,
                        rows_per_sample=A)
        v = self.buf("loss.v", (Me, ad), F32)
        ops.linear_f32(sufout, self.p("action_out_w"), v, Me, ad, D1, bias=self.p("action_out_b"))
        aloss = self.buf("loss.aloss", (B,), F32)
        dv = self.buf("loss.dv", (Me, ad), F32) if compute_grad_seed else None
        ops.mse_fwd_bwd(v, self._bufs["suf.u_t"], aloss, dv, B, A * ad, cfg.action_loss_weight / st.n_action)
        loss = self.buf("loss.total", (1,), F32)
        ops.weighted_sum(nll, st.ce_weights, loss, R, 1.0, False)
        ops.weighted_sum(aloss, None, loss, B, cfg.action_loss_weight / st.n_action, True)
        return loss, (X, XE)

    def _metrics(self, st: Staged) -> dict[str, torch.Tensor]:
        """Scalar metrics of lap.py:260,300,548-554,567 from the per-row nll / per-sample action loss (tiny arrays)."""
        nll, aloss = self.buf("loss.nll", (st.R,), F32), self.buf("loss.aloss", (st.B,), F32)
        per_sample = torch.zeros(st.B, dtype=F32, device=self.device)
        per_sample.index_add_(0, st.ce_sample, nll * st.ce_inv_count)
        m = {"lang_loss": per_sample.mean(), "action_loss": aloss.mean()}
        m["langact_loss"] = (per_sample * st.sample_mask).sum() / st.sample_mask.sum().clamp_min(1.0)
        return m

    def draw_augmentation(self, rng, observation: Observation, B: int) -> dict:
        """Augmentation parameters for every camera (model_adapter.py:118-151 draws them with jax.random per sample and
        camera; here numpy's generator seeded by rng).  VQA samples are skipped like the reference's `vqa_mask`."""
        gen = np.random.default_rng((int(rng) if rng is not None else 0) + 0x5EED)
        vqa = getattr(observation, "is_vqa_sample", None)
        skip = None if vqa is None else to_numpy(vqa).astype(np.float32)
        S = self.cfg.image_size
        return {k: image_tools.draw_augmentation_params(gen, B, S, S, skip) for k in self.cfg.image_keys}

    def compute_loss(self, rng, observation: Observation, actions, *, train: bool = False, stage_config=None,
                     verbose_mode=None, return_augmented_images: bool = False, noise=None, time=None, aug=None):
        """lap.py:380-602.  Returns (loss, metrics) — 0-d CUDA tensors.  `noise` [B,A,ad] and `time` [B] replace the
        reference's jax.random draws (lap.py:193-194); if omitted they are drawn from torch's generator seeded by rng.
        `aug` {camera: [B, 7] parameters} replaces the augmentation draws when `train` and the config enables image
        augmentation (model_adapter.py:118-151; parameter meaning: lapb200_image_augment)."""
        cfg = self.cfg
        B = to_numpy(actions).shape[0] if not isinstance(actions, torch.Tensor) else actions.shape[0]
        if train and cfg.enable_image_augmentation and aug is None:
            aug = self.draw_augmentation(rng, observation, B)
        if noise is None or time is None:
            gen = torch.Generator().manual_seed(int(rng) if rng is not None else 0)
            if noise is None:
                noise = torch.randn((B, cfg.action_horizon, cfg.action_dim), generator=gen)
            if time is None:  # Beta(1.5, 1) by inverse CDF, from the same seeded generator (lap.py:194)
                time = torch.rand((B,), generator=gen).pow(1.0 / 1.5) * 0.999 + 0.001
        st = self._stage(observation, actions, noise, time, with_loss=True, aug=aug if train else None)
        loss, _ = self._forward_loss(st, save=False, compute_grad_seed=False)
        metrics = self._metrics(st)
        if return_augmented_images:  # lap.py:413-420: the images the towers actually saw, for logging
            metrics = dict(metrics, augmented_images={
                k: (im.float() / 255.0 * 2.0 - 1.0 if im.dtype == torch.uint8 else im.clone())
                for k, im in zip(cfg.image_keys, st.images)})
        return loss[0].clone(), metrics

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def forward_backward(self, st: Staged, *, zero_grads: bool = True, softmax_mode: int = 0):
        """Forward + hand-written backward; fills self.G (flat fp32 grads).  Returns the loss (device scalar [1])."""
        loss = self.forward_backward_llm(st, zero_grads=zero_grads, softmax_mode=softmax_mode)
        self.backward_vision(st)
        return loss

    def llm_grad_range(self) -> tuple[int, int]:
        """Flat-buffer range whose gradients are final once forward_backward_llm returns (Gemma + expert + action
        projections + embedding): the data-parallel all-reduce of this range overlaps backward_vision."""
        return self.layout.offsets["g.qkv_w"], self.layout.small_begin

    def backward_vision(self, st: Staged) -> None:
        """SigLIP tower backward (the last ~12 % of the step); needs forward_backward_llm to have run."""
        self._siglip_bwd(st, self._bufs["bwd.dX"], self.cfg.prefix_len)

    def backward_vision_segment(self, st: Staged, l_hi: int, l_lo: int) -> None:
        """Encoder blocks l_hi-1 ... l_lo of the SigLIP backward; the first segment (l_hi = depth) also does the head,
        the last one (l_lo = 0) the patch embedding."""
        depth = self.cfg.siglip.depth
        if l_hi == depth:
            d = self._bwd_dims(st)
            self._siglip_bwd_head(st, self.buf("bwd.dX", (d["Mg"], d["D"])), self.cfg.prefix_len)
        self._siglip_bwd_layers(st, l_hi, l_lo)
        if l_lo == 0:
            self._siglip_bwd_tail(st)

    LLM_LAYER_GRADS = ("g.qkv_w", "g.o_w", "g.gu_w", "g.down_w", "e.qkv_w", "e.o_w", "e.gu_w", "e.down_w")
    VIS_LAYER_GRADS = ("img.qkv_w", "img.out_w", "img.fc1_w", "img.fc2_w")

    def layer_grad_ranges(self, names, l_lo: int, l_hi: int) -> list[tuple[int, int]]:
        """Flat-buffer ranges of the layers [l_lo, l_hi) of the per-layer tensors `names` (leading axis = layer)."""
        out = []
        for n in names:
            per = math.prod(self.layout.shapes[n][1:])
            o = self.layout.offsets[n]
            out.append((o + l_lo * per, o + l_hi * per))
        return out

    def forward_backward_llm(self, st: Staged, *, zero_grads: bool = True, softmax_mode: int = 0):
        """Forward of everything + backward through the loss heads, the transformer stack, the text embedding and
        the suffix embedding.  Leaves d(prefix tokens) in `bwd.dX` for backward_vision.  Equivalent to
        `forward_and_heads` -> `backward_llm_layers(depth, 0)` -> `backward_llm_tail` (the segments the data-parallel
        trainer interleaves with gradient all-reduces)."""
        loss = self.forward_and_heads(st, zero_grads=zero_grads, softmax_mode=softmax_mode)
        self.backward_llm_layers(st, self.cfg.gemma.depth, 0)
        self.backward_llm_tail(st)
        return loss

    def _bwd_dims(self, st: Staged):
        cfg, g, e = self.cfg, self.cfg.gemma, self.cfg.expert
        B, Pn, A = st.B, cfg.prefix_len, cfg.action_horizon
        hd, NH = g.head_dim, g.num_heads
        T = Pn + A
        return dict(B=B, Pn=Pn, A=A, L=cfg.max_token_len, C=len(cfg.image_keys), Np=cfg.num_patches, D=g.width,
                    D1=e.width, ad=cfg.action_dim, V=cfg.vocab_size, hd=hd, NH=NH, QKV=(NH + 2) * hd, F=g.mlp_dim,
                    F1=e.mlp_dim, T=T, Tpad=_round_up(T, 64), Mg=B * Pn, Me=B * A, R=st.R, Rq=T * NH, nm=P.n_mod(cfg))

    def forward_and_heads(self, st: Staged, *, zero_grads: bool = True, softmax_mode: int = 0):
        """Forward of everything + backward through the two loss heads (action projection, final norms, LM head).
        Leaves d(residual streams) of the last layer in `bwd.dX` / `bwd.dXE`."""
        assert self.G is not None, "allocate model.G (flat grads) first"
        cfg, g, e = self.cfg, self.cfg.gemma, self.cfg.expert
        lay = self.layout
        if zero_grads:
            self.G[lay.small_begin:].zero_()  # atomically-accumulated small tensors; GEMM wgrads overwrite theirs
        loss, (XL, XEL) = self._forward_loss(st, save=True, compute_grad_seed=True, softmax_mode=softmax_mode)
        d = self._bwd_dims(st)
        B, A, D, D1, ad, V, Mg, Me, R, nm = (d[k] for k in ("B", "A", "D", "D1", "ad", "V", "Mg", "Me", "R", "nm"))
        nm3 = nm * 3 * D1
        mod = self._bufs["suf.mod"]
        dmod = self.buf("bwd.dmod", (B, nm3), zero=True)
        dmod.zero_()
        bufs = self._bufs
        # ---- action head ----
        dv, sufout = bufs["loss.dv"], bufs["loss.sufout"]
        ops.sgemm(dv, sufout, self.g("action_out_w"), ad, D1, Me, 1, ad, 1, D1, ldc=D1)
        ops.sgemm(self.ones, dv, self.g("action_out_b"), 1, ad, Me, 0, 1, 1, ad, ldc=ad)
        dsuf = self.buf("bwd.dsuf", (Me, D1))
        ops.sgemm(dv, self.p("action_out_w"), dsuf, Me, D1, ad, ad, 1, 1, D1, ldc=D1)
        dXE = self.buf("bwd.dXE", (Me, D1))
        ops.ada_rmsnorm_bwd(dsuf, XEL, mod.view(-1)[(nm - 1) * 3 * D1:], nm3, bufs["loss.rstdEF"], None, dXE,
                            dmod.view(-1)[(nm - 1) * 3 * D1:], nm3, B, A, D1)
        # ---- LM head ----
        dlogits, pre2 = bufs["loss.dlogits"], bufs["loss.pre2"]
        dpre = self.buf("bwd.dpre", (R, D))
        self._dgrad(dlogits, self.w("g.embed"), dpre, R, V, D)
        self._wgrad(dlogits, pre2, self.g("g.embed"), R, V, D, ldx=2 * D)  # overwrites the whole table gradient
        dX = self.buf("bwd.dX", (Mg, D))
        dX.zero_()
        ops.rmsnorm_bwd(dpre, XL, self.p("g.final_norm_s"), bufs["loss.rstdF"], None, dX, self.g("g.final_norm_s"),
                        R, D, row_idx=st.ce_rows)
        return loss

    def backward_llm_layers(self, st: Staged, l_hi: int, l_lo: int) -> None:
        """Backward through transformer layers l_hi-1 ... l_lo (both experts + the shared attention).  When it
        returns, the weight gradients of those layers are final (the per-layer norm scales live in the small tail)."""
        cfg, g, e = self.cfg, self.cfg.gemma, self.cfg.expert
        d = self._bwd_dims(st)
        B, Pn, A, D, D1, hd, NH, QKV, F, F1, Tpad, Mg, Me, Rq, nm = (d[k] for k in (
            "B", "Pn", "A", "D", "D1", "hd", "NH", "QKV", "F", "F1", "Tpad", "Mg", "Me", "Rq", "nm"))
        nm3 = nm * 3 * D1
        bufs = self._bufs
        mod = self.buf("suf.mod", (B, nm3))
        dmod = self.buf("bwd.dmod", (B, nm3))
        dX = self.buf("bwd.dX", (Mg, D))
        dXE = self.buf("bwd.dXE", (Me, D1))
        dact = self.buf("bwd.dact", (Mg, F))
        dh = self.buf("bwd.dh", (Mg, D))
        dqkv0 = self.buf("bwd.dqkv0", (Mg, QKV))
        dyE = self.buf("bwd.dyE", (Me, D1))
        dactE = self.buf("bwd.dactE", (Me, F1))
        dhE = self.buf("bwd.dhE", (Me, D1))
        dqkv1 = self.buf("bwd.dqkv1", (Me, QKV))
        dOc = self.buf("bwd.dOc", (B, Rq, hd))
        dP = self.buf("bwd.dP", (B, Rq, Tpad))
        dQ = self.buf("bwd.dQ", (B, Rq, hd))
        dKc = self.buf("bwd.dKc", (B, Tpad, hd))
        dVc = self.buf("bwd.dVc", (B, Tpad, hd))
        positions = bufs["mask.pos"]
        qscale = hd ** -0.5
        for l in reversed(range(l_lo, l_hi)):
            sv = lambda n: bufs[f"g.{n}.{l}"]
            # ===== MLP, prefix expert =====
            GU, h2, X1 = sv("GU"), sv("h2"), sv("X1")
            # (fusing geglu_bwd into this GEMM's epilogue — EPI_GEGLU_BWD — measured slower: the epilogue becomes the
            #  bottleneck; the streaming kernel runs at the HBM roofline instead)
            self._dgrad(dX, self.w("g.down_w", l), dact, Mg, D, F)
            if self.save_mlp_act:
                ops.geglu_bwd(dact, GU, Mg, F, write_act=False)  # GU <- [dg|du]; act was kept by the forward
                self._wgrad(dX, sv("act"), self.g("g.down_w", l), Mg, D, F)
            else:
                ops.geglu_bwd(dact, GU, Mg, F)  # dact <- act, GU <- [dg|du]
                self._wgrad(dX, dact, self.g("g.down_w", l), Mg, D, F)
            self._wgrad(GU, h2, self.g("g.gu_w", l), Mg, 2 * F, D)
            self._dgrad(GU, self.w("g.gu_w", l), dh, Mg, 2 * F, D)
            ops.rmsnorm_bwd(dh, X1, self.p("g.ffn_norm_s", l), sv("rstd2"), dX, dX, self.g("g.ffn_norm_s", l), Mg, D)
            # ===== MLP, action expert =====
            GUE, hE2, XE1 = sv("GUE"), sv("hE2"), sv("XE1")
            ops.gated_bwd(dXE, sv("yEf"), mod.view(-1)[(2 * l + 1) * 3 * D1 + 2 * D1:], nm3, dyE,
                          dmod.view(-1)[(2 * l + 1) * 3 * D1 + 2 * D1:], nm3, B, A, D1)
            self._dgrad(dyE, self.w("e.down_w", l), dactE, Me, D1, F1)
            ops.geglu_bwd(dactE, GUE, Me, F1)
            self._wgrad(dyE, dactE, self.g("e.down_w", l), Me, D1, F1)
            self._wgrad(GUE, hE2, self.g("e.gu_w", l), Me, 2 * F1, D1)
            self._dgrad(GUE, self.w("e.gu_w", l), dhE, Me, 2 * F1, D1)
            ops.ada_rmsnorm_bwd(dhE, XE1, mod.view(-1)[(2 * l + 1) * 3 * D1:], nm3, sv("rstdE2"), dXE, dXE,
                                dmod.view(-1)[(2 * l + 1) * 3 * D1:], nm3, B, A, D1)
            # ===== attention output projections =====
            O0, O1 = sv("O0"), sv("O1")
            self._wgrad(dX, O0, self.g("g.o_w", l), Mg, D, NH * hd)
            # d(attention output), written by the two dgrad GEMMs straight into the stacked [B, (P+A)*NH, hd] layout
            # (one batch entry per sample: rows [0, P) from the prefix expert, [P, P+A) from the action expert)
            self._dgrad(dX, self.w("g.o_w", l), dOc, Pn, D, NH * hd, batch_i=B, a_bs=(Pn * D, 0), c_bs=(Rq * hd, 0))
            ops.gated_bwd(dXE, sv("yEa"), mod.view(-1)[(2 * l) * 3 * D1 + 2 * D1:], nm3, dyE,
                          dmod.view(-1)[(2 * l) * 3 * D1 + 2 * D1:], nm3, B, A, D1)
            self._wgrad(dyE, O1, self.g("e.o_w", l), Me, D1, NH * hd)
            self._dgrad(dyE, self.w("e.o_w", l), dOc.view(-1)[Pn * NH * hd:], A, D1, NH * hd, batch_i=B,
                        a_bs=(A * D1, 0), c_bs=(Rq * hd, 0))
            # ===== shared attention =====
            Pm, Q, Kc, Vc = sv("P"), sv("Q"), sv("Kc"), sv("Vc")
            bsR, bsT, bsP = (Rq * hd, 0), (Tpad * hd, 0), (Rq * Tpad, 0)
            if self.fuse_softmax_bwd:
                # dS = P o (dP - rowsum(P o dP)) with rowsum(P o dP) = dO . O: the row term is one small kernel over
                # (dO, O) and the dP GEMM's epilogue emits dS directly — no separate pass over P and dP
                delta = self.buf("bwd.delta", (B, Rq), F32)
                ops.rowdot(dOc, O0, delta, Pn * NH, hd, hd, hd, nbi=B, d_bs=(Rq * hd, 0), o_bs=(Pn * NH * hd, 0),
                           out_rows=Rq, out_off=0)
                ops.rowdot(dOc.view(-1)[Pn * NH * hd:], O1, delta, A * NH, hd, hd, hd, nbi=B, d_bs=(Rq * hd, 0),
                           o_bs=(A * NH * hd, 0), out_rows=Rq, out_off=Pn * NH)
                ops.gemm(dOc, Vc, dP, M=Rq, N=Tpad, K=hd, ldc=Tpad, batch_i=B, a_bs=bsR, b_bs=bsT, c_bs=bsP,
                         epi=ops.EPI_SOFTMAX_BWD, C2=Pm, ldc2=Tpad, bias=delta)
            else:
                ops.gemm(dOc, Vc, dP, M=Rq, N=Tpad, K=hd, ldc=Tpad, batch_i=B, a_bs=bsR, b_bs=bsT, c_bs=bsP)
            if not cfg.stop_action_to_vlm_grad:
                ops.gemm(Pm, dOc, dVc, M=Tpad, N=hd, K=Rq, a_major=1, b_major=1, lda=Tpad, ldb=hd, ldc=hd, batch_i=B,
                         a_bs=bsP, b_bs=bsR, c_bs=bsT)
            if not self.fuse_softmax_bwd:
                ops.softmax_bwd(Pm, dP, dP, B * Rq, Tpad)
            ops.gemm(dP, Kc, dQ, M=Rq, N=hd, K=Tpad, b_major=1, lda=Tpad, ldb=hd, ldc=hd, batch_i=B, a_bs=bsP,
                     b_bs=bsT, c_bs=bsR)
            if cfg.stop_action_to_vlm_grad:
                # gemma.py:206-213,242-269: action-expert queries read expert 0's K and V through stop_gradient, so the
                # [action rows x prefix keys] block of P and dS must not reach dV / dK (dQ above used the full dS).
                # Both tensors are dead after this layer's two products: the block is cleared in place.
                Pm.view(B, Rq, Tpad)[:, Pn * NH:, :Pn].zero_()
                dP.view(B, Rq, Tpad)[:, Pn * NH:, :Pn].zero_()
                ops.gemm(Pm, dOc, dVc, M=Tpad, N=hd, K=Rq, a_major=1, b_major=1, lda=Tpad, ldb=hd, ldc=hd, batch_i=B,
                         a_bs=bsP, b_bs=bsR, c_bs=bsT)
            ops.gemm(dP, Q, dKc, M=Tpad, N=hd, K=Rq, a_major=1, b_major=1, lda=Tpad, ldb=hd, ldc=hd, batch_i=B,
                     a_bs=bsP, b_bs=bsR, c_bs=bsT)
            ops.rope_bwd(dQ, dKc, dVc, positions, self.timescale, dqkv0, dqkv1, B, Pn, A, Tpad, NH, hd, qscale)
            # ===== QKV projections + pre-attention norms =====
            h, X = sv("h"), sv("X")
            self._wgrad(dqkv0, h, self.g("g.qkv_w", l), Mg, QKV, D)
            self._dgrad(dqkv0, self.w("g.qkv_w", l), dh, Mg, QKV, D)
            ops.rmsnorm_bwd(dh, X, self.p("g.attn_norm_s", l), sv("rstd"), dX, dX, self.g("g.attn_norm_s", l), Mg, D)
            hE, XE = sv("hE"), sv("XE")
            self._wgrad(dqkv1, hE, self.g("e.qkv_w", l), Me, QKV, D1)
            self._dgrad(dqkv1, self.w("e.qkv_w", l), dhE, Me, QKV, D1)
            ops.ada_rmsnorm_bwd(dhE, XE, mod.view(-1)[(2 * l) * 3 * D1:], nm3, sv("rstdE"), dXE, dXE,
                                dmod.view(-1)[(2 * l) * 3 * D1:], nm3, B, A, D1)

    def backward_llm_tail(self, st: Staged) -> None:
        """Text-embedding scatter, adaRMS modulation Dense, time MLP and action_in_proj gradients."""
        cfg = self.cfg
        d = self._bwd_dims(st)
        B, Pn, L, C, Np, D, D1, ad, Mg, Me, nm = (d[k] for k in ("B", "Pn", "L", "C", "Np", "D", "D1", "ad", "Mg", "Me", "nm"))
        nm3 = nm * 3 * D1
        bufs = self._bufs
        dmod = self.buf("bwd.dmod", (B, nm3))
        dX = self.buf("bwd.dX", (Mg, D))
        dXE = self.buf("bwd.dXE", (Me, D1))
        # ---- text embedding (scatter-add on top of the LM-head table gradient) ----
        ops.embed_bwd(st.tokens, dX, self.g("g.embed"), B, L, C * Np, Pn, D, math.sqrt(D))
        # ---- adaRMS modulation Dense + time MLP + action_in_proj ----
        cond16 = bufs["suf.cond16"]
        self._wgrad(dmod, cond16, self.g("e.mod_w").view(nm3, D1), B, nm3, D1)
        ops.colsum(dmod, nm3, self.g("e.mod_b").view(-1), B, nm3)
        dcond16 = self.buf("bwd.dcond16", (B, D1))
        self._dgrad(dmod, self.w("e.mod_w").view(nm3, D1), dcond16, B, nm3, D1)
        z1, s1, z2, te = bufs["suf.z1"], bufs["suf.s1"], bufs["suf.z2"], bufs["suf.te"]
        dz2, ds1, dz1 = (self.buf(f"bwd.{n}", (B, D1), F32) for n in ("dz2", "ds1", "dz1"))
        ops.swish_bwd(z2, None, dcond16, dz2, B * D1)
        ops.sgemm(dz2, s1, self.g("time_out_w"), D1, D1, B, 1, D1, 1, D1, ldc=D1)
        ops.sgemm(self.ones, dz2, self.g("time_out_b"), 1, D1, B, 0, 1, 1, D1, ldc=D1)
        ops.sgemm(dz2, self.p("time_out_w"), ds1, B, D1, D1, D1, 1, 1, D1, ldc=D1)
        ops.swish_bwd(z1, ds1, None, dz1, B * D1)
        ops.sgemm(dz1, te, self.g("time_in_w"), D1, D1, B, 1, D1, 1, D1, ldc=D1)
        ops.sgemm(self.ones, dz1, self.g("time_in_b"), 1, D1, B, 0, 1, 1, D1, ldc=D1)
        x_t = bufs["suf.x_t"]
        ops.sgemm(dXE, x_t, self.g("action_in_w"), D1, ad, Me, 1, D1, 1, ad, ldc=ad)
        ops.colsum(dXE, D1, self.g("action_in_b"), Me, D1)

    # ------------------------------------------------------------------------------------------
    # inference: prefix pass -> KV cache -> Euler steps  (lap.py:605-675)
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample_actions(self, rng, observation: Observation, *, num_steps: int = 10, noise=None) -> torch.Tensor:
        """lap.py:605-675.  After one eager call per (batch, num_steps) the whole prefix pass + Euler loop (~2000 kernel
        launches) is captured as ONE CUDA graph and replayed; inputs/outputs go through persistent device buffers."""
        cfg = self.cfg
        st = self._stage(observation, with_loss=False)
        B, A, ad = st.B, cfg.action_horizon, cfg.action_dim
        if noise is None:
            gen = torch.Generator().manual_seed(int(rng) if rng is not None else 0)
            noise = torch.randn((B, A, ad), generator=gen)
        noise_t = noise if isinstance(noise, torch.Tensor) else torch.from_numpy(np.asarray(noise))
        x = self.buf("inf.x", (B, A * ad), F32)
        x.copy_(noise_t.to(torch.float32).reshape(B, A * ad), non_blocking=True)
        key = (B, int(num_steps))
        g = self._infer_graphs.get(key)
        if g is None and self.use_cuda_graph and self._infer_warm.get(key, 0) >= 1:
            torch.cuda.synchronize()
            x_keep = x.clone()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._sample_actions_impl(st, num_steps)
            self._infer_graphs[key] = g
            x.copy_(x_keep)
        if g is None:
            self._sample_actions_impl(st, num_steps)
            self._infer_warm[key] = self._infer_warm.get(key, 0) + 1
        else:
            g.replay()
        return x.view(B, A, ad).clone()

    def _sample_actions_impl(self, st: Staged, num_steps: int) -> None:
        cfg, g, e = self.cfg, self.cfg.gemma, self.cfg.expert
        B, Pn, A, L = st.B, cfg.prefix_len, cfg.action_horizon, cfg.max_token_len
        C, Np = len(cfg.image_keys), cfg.num_patches
        D, D1, ad = g.width, e.width, cfg.action_dim
        T = Pn + A
        Tpad = _round_up(T, 64)
        W32 = Tpad // 32
        x = self.buf("inf.x", (B, A * ad), F32)
        # ---- prefix pass fills the cache ----
        X0 = self.buf("inf.X0", (B * Pn, D))
        self._siglip_fwd(st, X0, Pn, save=False)
        ops.embed_fwd(st.tokens, self.p("g.embed"), X0, B, L, C * Np, Pn, D, math.sqrt(D))
        bits_p = self.buf("inf.bits_p", (B, Pn, W32), torch.int32)
        pos_p = self.buf("inf.pos_p", (B, Pn), torch.int32)
        ops.mask_build(st.pm, st.par, None, None, None, bits_p, pos_p, B, Pn, 0, W32)
        Kc = self.buf("inf.Kc", (g.depth, B, Tpad, g.head_dim), zero=True)
        Vc = self.buf("inf.Vc", (g.depth, B, Tpad, g.head_dim), zero=True)
        self._gemma_prefix_only(B, X0, bits_p, pos_p, (Kc, Vc))
        # ---- suffix rows: mask [B, A, P+A] and positions (lap.py:641-654; prefix part = prefix_mask alone) ----
        bits_s = self.buf("inf.bits_s", (B, A, W32), torch.int32)
        pos_s = self.buf("inf.pos_s", (B, A), torch.int32)
        ops.mask_build(st.pm, st.par, st.pm, st.sm, st.sar, bits_s, pos_s, B, Pn, A, W32, row_begin=Pn, infer_rows=True)
        dt = -1.0 / num_steps
        if self.use_denoise_megakernel and ops.denoise_supported(B, A, ad, D1, e.num_heads, e.head_dim, e.mlp_dim, Pn, Tpad,
                                                                 num_steps):
            self._denoise_loop_fused(x, Kc, Vc, bits_s, pos_s, num_steps, dt)
            return
        t = 1.0
        tbuf = self.buf("inf.t", (B,), F32)
        XE0 = self.buf("inf.XE0", (B * A, D1))
        sufout = self.buf("inf.sufout", (B * A, D1))
        rstd = self.buf("inf.rstd", (B * A,), F32)
        v = self.buf("inf.v", (B * A, ad), F32)
        nm = P.n_mod(cfg)
        nm3 = nm * 3 * D1
        n_iter = 0
        while t >= -dt / 2:  # lap.py:669-672: exactly num_steps iterations
            tbuf.fill_(t)
            self._suffix_embed(st, x, tbuf, XE0)
            _, XE = self._gemma_fwd(B, None, XE0, bits_s, pos_s, save=False, kv_cache=(Kc, Vc), tag="inf")
            mod = self._bufs["suf.mod"]
            ops.rmsnorm_fwd(XE, sufout, rstd, B * A, D1, mod=mod.view(-1)[(nm - 1) * 3 * D1:], ldmod=nm3,
                            rows_per_sample=A)
            ops.linear_f32(sufout, self.p("action_out_w"), v, B * A, ad, D1, bias=self.p("action_out_b"))
            ops.axpy(x, v, dt, B * A * ad)
            t += dt
            n_iter += 1
        assert n_iter == num_steps

    def _denoise_loop_fused(self, x, Kc, Vc, bits_s, pos_s, num_steps: int, dt: float) -> None:
        """K10 (csrc/denoise.cu): all Euler steps of lap.py:634-672 for one sample in ONE persistent cooperative kernel.
        The prefix values are transposed once ([L, hd, keys]) so that P·V reads keys contiguously."""
        cfg, g, e = self.cfg, self.cfg.gemma, self.cfg.expert
        Pn, A, ad, D1, L = cfg.prefix_len, cfg.action_horizon, cfg.action_dim, e.width, g.depth
        NH, HD, F1 = e.num_heads, e.head_dim, e.mlp_dim
        Tpad = Kc.shape[2]
        TpadK = _round_up(Pn, 64)
        nm = P.n_mod(cfg)
        VcT = self.buf("inf.VcT", (L, HD, TpadK))
        packed = self.denoise_cluster
        if packed:
            # tile-major copies of the prefix keys and of V^T (once per inference, like the transpose it replaces)
            Kp = self.buf("inf.Kp", (L, Tpad, HD))
            ops.pack_tiles(Kc, Kp, Tpad, HD, HD, 1, HD, batch=L, src_bs=Kc.shape[1] * Tpad * HD, dst_bs=Tpad * HD)
            ops.pack_tiles(Vc, VcT, HD, TpadK, 1, HD, Pn, batch=L, src_bs=Vc.shape[1] * Tpad * HD, dst_bs=HD * TpadK)
            Kc = Kp
            pw = self._packed_expert_weights()
        else:
            ops.transpose_v(Vc, VcT, L, Tpad, TpadK, HD, Pn)
            pw = None
        wsel = (lambda n: pw[n][0]) if packed else (lambda n: self.w(n, 0))
        nch = TpadK // 64 + 1
        times, t = [], 1.0
        while t >= -dt / 2:  # lap.py:669-672
            times.append(t)
            t += dt
        assert len(times) == num_steps
        S = num_steps
        ptrs = dict(
            qkv_w=wsel("e.qkv_w"), o_w=wsel("e.o_w"), gu_w=wsel("e.gu_w"), down_w=wsel("e.down_w"),
            mod_w=self.w("e.mod_w"), mod_b=self.p("e.mod_b"),
            ain_w=self.p("action_in_w"), ain_b=self.p("action_in_b"), tin_w=self.p("time_in_w"), tin_b=self.p("time_in_b"),
            tout_w=self.p("time_out_w"), tout_b=self.p("time_out_b"), aout_w=self.p("action_out_w"),
            aout_b=self.p("action_out_b"), Kc=Kc, VcT=VcT, bits=bits_s, pos=pos_s, timescale=self.timescale, x=x,
            s1=self.buf("dn.s1", (S, D1), F32), cond16=self.buf("dn.cond16", (S, D1)),
            mod=self.buf("dn.mod", (S, nm * 3 * D1)), XE=self.buf("dn.XE", (16, D1)), XE1=self.buf("dn.XE1", (16, D1)),
            qkv=self.buf("dn.qkv", (16, (NH + 2) * HD)), O=self.buf("dn.O", (16, NH * HD)), act=self.buf("dn.act", (16, F1)),
            part_o=self.buf("dn.part_o", (NH * nch, 16, HD), F32), part_ml=self.buf("dn.part_ml", (NH * nch, 16, 2), F32),
            sync=self.buf("dn.sync", (32,), torch.int32, zero=True))  # LAPB_DENOISE_SYNC_WORDS
        if self.denoise_profile:
            ptrs["prof"] = self.buf("dn.prof", (256 * 32,), torch.int64, zero=True)  # [CTA][slot]

        def lstride(name):
            return self.w(name, 1).data_ptr() - self.w(name, 0).data_ptr() >> 1 if L > 1 else 0

        ops.denoise_loop(
            ints=dict(A=A, ad=ad, D1=D1, NH=NH, HD=HD, F1=F1, L=L, Pn=Pn, Tpad=Tpad, TpadK=TpadK, W32=Tpad // 32, nm=nm,
                      num_steps=S, packed=int(packed), flags=int(os.environ.get("LAPB_DENOISE_FLAGS", "0"))),
            dt=dt, qscale=HD ** -0.5, times=times, ptrs=ptrs,
            strides=dict(qkv_ls=lstride("e.qkv_w"), o_ls=lstride("e.o_w"), gu_ls=lstride("e.gu_w"),
                         down_ls=lstride("e.down_w"), kc_ls=Tpad * HD, vct_ls=HD * TpadK))

    def denoise_error_flag(self) -> int:
        """1 if a grid barrier of the last fused denoise loop timed out (never expected; checked by the tests)."""
        b = self._bufs.get("dn.sync")
        return int(b[1].item()) if b is not None else 0

    # ------------------------------------------------------------------------------------------
    # autoregressive decoding of lang-action tokens (lap.py:678-766), greedy
    # ------------------------------------------------------------------------------------------
    def sample_tokens(self, rng, observation: Observation, *, max_decoding_steps: int = 390,
                      temperature: float = 0.0, gumbel=None) -> torch.Tensor:
        """LAP.sample_tokens: prefix prefill -> KV cache -> one token per step through expert 0 alone, until every
        sample has emitted EOS or `max_decoding_steps`.  Returns int32 [B, max_decoding_steps] (zeros after the stop).

        The reference right-aligns the prefix (pi0_fast.left_to_right_align) and masks decode steps by slot RANGE
        (slot >= prefix_start, lap.py:737-741).  Attention does not depend on where a key is stored, so the engine keeps
        its left-aligned cache and translates the range into key positions: rolled slot i holds token (i + seqlen) mod P."""
        # temperature > 0 (lap.py:727-729): jax.random.categorical(key, z) == argmax(z + Gumbel(key)).  The Gumbel noise is
        # either passed in (`gumbel` [B, max_decoding_steps, vocab], what the parity tests do) or drawn on the device from
        # a torch generator seeded by `rng` — same distribution, not JAX's threefry stream.
        cfg, g = self.cfg, self.cfg.gemma
        temperature = float(temperature)
        gum_gen = None
        if temperature > 0.0:
            if gumbel is not None:
                gumbel = (gumbel if isinstance(gumbel, torch.Tensor) else torch.from_numpy(np.asarray(gumbel))).to(
                    self.device, torch.float32)
            else:
                gum_gen = torch.Generator(device=self.device).manual_seed(int(rng) if rng is not None else 0)
        st = self._stage(observation, with_loss=False)
        B, Pn, L, S = st.B, cfg.prefix_len, cfg.max_token_len, int(max_decoding_steps)
        if B > 16:
            raise NotImplementedError("sample_tokens uses the weight-streaming kernels: batch <= 16")
        C, Np, D, hd, NH, F = len(cfg.image_keys), cfg.num_patches, g.width, g.head_dim, g.num_heads, g.mlp_dim
        QKV = (NH + 2) * hd
        V = cfg.vocab_size
        # ---- prefill (same kernels as sample_actions' prefix pass, but the last layer's output is needed) ----
        Tpad = _round_up(Pn + cfg.action_horizon, 64)
        W32 = Tpad // 32
        X0 = self.buf("inf.X0", (B * Pn, D))
        self._siglip_fwd(st, X0, Pn, save=False)
        ops.embed_fwd(st.tokens, self.p("g.embed"), X0, B, L, C * Np, Pn, D, math.sqrt(D))
        bits_p = self.buf("inf.bits_p", (B, Pn, W32), torch.int32)
        pos_p = self.buf("inf.pos_p", (B, Pn), torch.int32)
        ops.mask_build(st.pm, st.par, None, None, None, bits_p, pos_p, B, Pn, 0, W32)
        Kp = self.buf("inf.Kc", (g.depth, B, Tpad, hd), zero=True)
        Vp = self.buf("inf.Vc", (g.depth, B, Tpad, hd), zero=True)
        X = self._gemma_fwd_prefix(B, X0, bits_p, pos_p, (Kp, Vp), need_output=True)
        # ---- decode cache: prefix keys + one slot per step ----
        Tcap = _round_up(Pn + S, 64)
        W32c = Tcap // 32
        Kc = self.buf("ar.Kc", (g.depth, B, Tcap, hd), zero=True)
        Vc = self.buf("ar.Vc", (g.depth, B, Tcap, hd), zero=True)
        Kc[:, :, :Pn].copy_(Kp[:, :, :Pn])
        Vc[:, :, :Pn].copy_(Vp[:, :, :Pn])
        pm = st.pm.cpu().numpy().astype(bool)  # [B, Pn] validity of the prefix tokens (tiny, host side)
        idx = np.arange(Pn)
        seqlen = (pm * idx[None, :]).max(-1) + 1                      # pi0_fast.py:60
        prefill_len = pm.sum(-1)
        prefix_start = Pn - prefill_len
        rolled_slot = (idx[None, :] - seqlen[:, None]) % Pn            # slot of token j after the roll by -seqlen
        allowed = np.zeros((B, Tcap), dtype=bool)
        allowed[:, :Pn] = rolled_slot >= prefix_start[:, None]
        allowed[:, Pn:] = True                                         # decode slots; S_len cuts off the future ones
        bits_np = np.packbits(allowed.reshape(B, W32c, 32), axis=-1, bitorder="little").view(np.uint32).reshape(B, 1, W32c)
        bits_ar = torch.from_numpy(bits_np.view(np.int32).copy()).to(self.device)
        pos0 = torch.from_numpy(prefill_len.astype(np.int32)).to(self.device)
        last_row = torch.from_numpy((np.arange(B) * Pn + seqlen - 1).astype(np.int64)).to(self.device)
        # ---- first logits: final norm + LM head on the last valid prefix token ----
        pre2 = self.buf("ar.pre2", (B, 2 * D))
        rstd = self.buf("ar.rstd", (B,), F32)
        logits = self.buf("ar.logits", (B, V), F32)
        ops.rmsnorm_fwd(X, pre2, rstd, B, D, scale=self.p("g.final_norm_s"), row_idx=last_row, ldy=2 * D, dup=True)
        ops.skinny_gemm(pre2, self.E_split, logits, M=B, N=V, K=2 * D)
        out = torch.zeros((B, S), dtype=torch.int32, device=self.device)
        eos = torch.zeros((B,), dtype=torch.bool, device=self.device)
        x = self.buf("ar.x", (B, D))
        x1 = self.buf("ar.x1", (B, D))
        h = self.buf("ar.h", (B, D))
        qkv = self.buf("ar.qkv", (B, QKV))
        Q = self.buf("ar.Q", (B, 1, NH, hd))
        O = self.buf("ar.O", (B, NH * hd))
        act = self.buf("ar.act", (B, F))
        pos = self.buf("ar.pos", (B,), torch.int32)
        qscale = hd ** -0.5
        step = 0
        while step < S:
            if temperature > 0.0:
                if gumbel is not None:
                    gstep = gumbel[:, step]
                else:  # -log(-log(u)), u in (0, 1)
                    u = torch.rand((B, V), generator=gum_gen, device=self.device).clamp_(1e-20, 1.0 - 1e-7)
                    gstep = -torch.log(-torch.log(u))
                token = torch.argmax(logits / temperature + gstep, dim=-1).to(torch.int32)   # lap.py:727-729
            else:
                token = torch.argmax(logits, dim=-1).to(torch.int32)   # lap.py:730 (temperature = 0)
            out[:, step] = token
            eos |= token == self.EOS_TOKEN
            all_eos = bool(eos.all().item())
            # decode one step with expert 0 (the reference also runs it after the last token; its result is unused)
            if all_eos or step + 1 >= S:
                break
            ops.embed_fwd(token, self.p("g.embed"), x, B, 1, 0, 1, D, math.sqrt(D))
            torch.add(pos0, step, out=pos)
            xin = x
            for l in range(g.depth):
                ops.rmsnorm_fwd(xin, h, rstd, B, D, scale=self.p("g.attn_norm_s", l))
                ops.skinny_gemm(h, self.w("g.qkv_w", l), qkv, M=B, N=QKV, K=D)
                ops.rope_fwd(None, qkv, pos, self.timescale, Q, Kc[l], Vc[l], B, Pn + step, 1, Tcap, NH, hd, Pn + step,
                             qscale)
                ops.decode_attn(Q, Kc[l], Vc[l], bits_ar, O, B, 1, NH, hd, Pn + step + 1, Tcap, W32c)
                ops.skinny_gemm(O, self.w("g.o_w", l), x1, M=B, N=D, K=NH * hd, epi=ops.EPI_RESID, resid=xin)
                ops.rmsnorm_fwd(x1, h, rstd, B, D, scale=self.p("g.ffn_norm_s", l))
                ops.skinny_gemm(h, self.w("g.gu_w", l), act, M=B, N=F, K=D, epi=ops.EPI_GEGLU)
                ops.skinny_gemm(act, self.w("g.down_w", l), x, M=B, N=D, K=F, epi=ops.EPI_RESID, resid=x1)
                xin = x
            ops.rmsnorm_fwd(x, pre2, rstd, B, D, scale=self.p("g.final_norm_s"), ldy=2 * D, dup=True)
            ops.skinny_gemm(pre2, self.E_split, logits, M=B, N=V, K=2 * D)
            step += 1
        return out

    def _gemma_prefix_only(self, B, X0, bits, positions, cache):
        """Prefix-only pass (lap.py:627): expert 0 alone, K/V (post-RoPE) written into the cache."""
        cfg = self.cfg
        # reuse the generic layer loop with A = 0 by temporarily viewing the sequence as prefix-only
        self._gemma_fwd_prefix(B, X0, bits, positions, cache)

    def _gemma_fwd_prefix(self, B, X, bits, positions, cache, need_output: bool = False):
        cfg, g = self.cfg, self.cfg.gemma
        Pn = cfg.prefix_len
        Tpad = _round_up(cfg.prefix_len + cfg.action_horizon, 64)
        W32 = Tpad // 32
        D, hd, NH, F = g.width, g.head_dim, g.num_heads, g.mlp_dim
        QKV = (NH + 2) * hd
        Mg = B * Pn
        R = Pn * NH
        qscale = hd ** -0.5
        # batch 1: the down projection (M = 692 rows, K = 16384) as deterministic split-K slabs over 2-CTA tiles, summed
        # (+ residual, same rounding points) inside the next layer's RMSNorm kernel (ops.resid_norm_fwd)
        ksd = self._slab_splits(F, 3) if (Mg <= 1024 and F >= 4096 and self.use_small_m_split_k) else 0
        accd = self.buf("inf.accd", (max(ksd, 1), Mg, D), F32) if ksd else None
        pending = None  # residual X1 of a down projection whose slabs still wait in accd
        for l in range(g.depth):
            Kc, Vc = cache[0][l], cache[1][l]
            h = self.buf("inf.h", (Mg, D))
            rstd = self.buf("inf.rstdp", (Mg,), F32)
            if pending is not None:
                X = self.buf(f"inf.Xp{l % 2}", (Mg, D))
                ops.resid_norm_fwd(pending, accd, ksd, Mg * D, None, X, False, self.p("g.attn_norm_s", l), None, h, None,
                                   rstd, Mg, D)
                pending = None
            else:
                ops.rmsnorm_fwd(X, h, rstd, Mg, D, scale=self.p("g.attn_norm_s", l))
            qkv0 = self.buf("inf.qkv0", (Mg, QKV))
            ops.gemm(h, self.w("g.qkv_w", l), qkv0, M=Mg, N=QKV, K=D)
            Q = self.buf("inf.Qp", (B, Pn, NH, hd))
            ops.rope_fwd(qkv0, None, positions, self.timescale, Q, Kc, Vc, B, Pn, 0, Tpad, NH, hd, 0, qscale)
            if l == g.depth - 1 and not need_output:
                # sample_actions consumes only the KV cache of the prefix pass (lap.py:627 discards the outputs): the
                # last layer's attention, output projection and MLP feed nothing
                return None
            O0 = self.buf("inf.O0", (Mg, NH * hd))
            if hd == 256 and self.use_fused_attention:
                ops.fa_gemma_fwd(Q, Kc, Vc, bits, None, O0, O0, B, R, NH, Pn, Pn, Tpad, W32, R, hd)
            else:
                S = self.buf("inf.Sp", (B, R, Tpad), F32)
                ops.gemm(Q, Kc, S, M=R, N=Tpad, K=hd, ldc=Tpad, batch_i=B, a_bs=(R * hd, 0), b_bs=(Tpad * hd, 0),
                         c_bs=(R * Tpad, 0))
                Pm = self.buf("inf.Pp", (B, R, Tpad))
                ops.attn_softmax_fwd(S, bits, Pm, B, R, NH, Pn, Tpad, W32)
                ops.gemm(Pm, Vc, O0, M=R, N=hd, K=Tpad, b_major=1, lda=Tpad, ldb=hd, ldc=hd, batch_i=B,
                         a_bs=(R * Tpad, 0), b_bs=(Tpad * hd, 0), c_bs=(R * hd, 0))
            X1 = self.buf("inf.X1", (Mg, D))
            ops.gemm(O0, self.w("g.o_w", l), X1, M=Mg, N=D, K=NH * hd, epi=ops.EPI_RESID, resid=X)
            h2 = self.buf("inf.h2", (Mg, D))
            ops.rmsnorm_fwd(X1, h2, rstd, Mg, D, scale=self.p("g.ffn_norm_s", l))
            act = self.buf("inf.act", (Mg, F))
            ops.gemm(h2, self.w("g.gu_w", l), act, M=Mg, N=F, K=D, epi=ops.EPI_GEGLU, C2=None)
            if ksd:
                ops.gemm(act, self.w("g.down_w", l), accd, M=Mg, N=D, K=F, k_splits=ksd, split_stride=Mg * D, cta_group=2,
                         block_n=256)
                pending = X1
            else:
                X2 = self.buf(f"inf.Xp{(l + 1) % 2}", (Mg, D))
                ops.gemm(act, self.w("g.down_w", l), X2, M=Mg, N=D, K=F, epi=ops.EPI_RESID, resid=X1)
                X = X2
        if pending is not None:  # the caller wants the last layer's output: finalise without a norm
            X = self.buf(f"inf.Xp{g.depth % 2}", (Mg, D))
            ops.resid_norm_fwd(pending, accd, ksd, Mg * D, None, X, False, None, None, None, None, None, Mg, D)
        return X
