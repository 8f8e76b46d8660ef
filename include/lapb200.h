/* lapb200.h — C ABI of liblapb200.so, the sm_100a kernel library behind lap_b200.
 *
 * The reference (lihzha/lap) is pure Python/JAX and has NO native interface: every
 * device op there is whatever XLA emits for the Flax modules.  This header is the
 * boundary the B200 engine crosses instead: plain pointers + sizes, no torch types.
 * Each entry point cites the reference op (file:line under /root/reference) whose
 * arithmetic it replaces; `OP/` = third_party/openpi/src/openpi/.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless named h_*; `stream` is a cudaStream_t.
 *   - bf16 = raw __nv_bfloat16 bits (uint16), row-major; ld* are in ELEMENTS.
 *   - return value: 0 on success, otherwise a cudaError_t (or -1 for bad arguments);
 *     lapb200_last_error() returns a static description.  Kernels never abort.
 *   - there is no CPU fallback: on a non-sm_100 device every call returns an error.
 */
#ifndef LAPB200_H_
#define LAPB200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* lapb_stream_t; /* cudaStream_t */

const char* lapb200_last_error(void);
int lapb200_version(void);
/* 0 if the current device is sm_100 (B200); otherwise an error code. */
int lapb200_check_device(void);

/* ------------------------------------------------------------------------- */
/* K3: tcgen05 bf16 GEMM (TMA-fed, TMEM accumulators), fused epilogues.      */
/*   C[b][M,N] = epi( A[b][M,K] * B[b][N,K]^T )                              */
/* Replaces jnp.einsum / jnp.dot / nn.Dense GEMMs:                           */
/*   OP/models/lora.py:57 (Einsum), :145 (FeedForward._dot),                 */
/*   src/lap/models/backbones/gemma.py:186-201,285 (q/kv/out einsums),       */
/*   OP/models/siglip.py:69-72,88-93,286 (Dense, MHA projections, head),     */
/*   and their autodiff transposes (dgrad/wgrad).                            */
/* ------------------------------------------------------------------------- */
enum {
  LAPB_EPI_NONE = 0,          /* C = acc (+bias); bf16 or fp32 out; optional C += (fp32)            */
  LAPB_EPI_BIAS_GELU = 1,     /* C2 = pre = bf16(bf16(acc)+bias); C = bf16(gelu_tanh(pre))  siglip.py:69-72 */
  LAPB_EPI_RESID = 2,         /* C = bf16(resid + y), y = bf16(acc) (+bias)                 gemma.py:582     */
  LAPB_EPI_GATED_RESID = 3,   /* C = bf16(resid + bf16(y*gate[row/gate_rows]))              gemma.py:583     */
  LAPB_EPI_GEGLU = 4,         /* dual-B: C=bf16(gelu(g)*u), C2[:, n]=g, C2[:, n+N]=u        lora.py:124-142  */
  LAPB_EPI_QSCALE = 5,        /* as NONE(+bias); columns < q_cols are divided by q_div      flax MHA q/sqrt(d) */
  LAPB_EPI_GEGLU_BWD = 6,     /* acc = dAct; C2 = [g|u] in, [dg|du] out (in place); C = act (recomputed)   lora.py:124-142 bwd */
  LAPB_EPI_GELU_BWD = 7,      /* C = bf16(acc) * gelu'(C2)                                   siglip.py:71 bwd */
  LAPB_EPI_SOFTMAX_BWD = 8    /* acc = dP; C = dS = C2 o (bf16(acc) - bias[batch*M + row]): C2 = P, bias = rowsum(dO o O)
                                 (lapb200_rowdot), batch = batch_o index * batch_i + batch_i index      gemma.py:261 bwd */
};

typedef struct {
  /* operands (bf16). major: 0 = K-major (A is [M,K], B is [N,K]); 1 = MN-major (A is [K,M], B is [K,N]) */
  const void* A;
  const void* B;
  int32_t a_major, b_major;
  int64_t lda, ldb;
  int64_t a_bs_i, a_bs_o, b_bs_i, b_bs_o; /* batch strides (elements): inner, outer */
  int32_t M, N, K;
  int32_t batch_i, batch_o;
  /* output */
  void* C;
  int64_t ldc, c_bs_i, c_bs_o;
  int32_t c_fp32;     /* 1: C is float, 0: bf16 */
  int32_t accumulate; /* 1: C += result (fp32 C only) */
  int32_t epi;
  /* epilogue operands */
  const float* bias; /* [N] fp32 or NULL */
  const void* resid; /* bf16, same indexing as C via ldr/r_bs_* */
  int64_t ldr, r_bs_i, r_bs_o;
  const void* gate; /* bf16 [rows/gate_rows, ldg] */
  int64_t ldg;
  int32_t gate_rows;
  void* C2; /* second output (bf16): pre-activation (BIAS_GELU) or [g|u] buffer (GEGLU) */
  int64_t ldc2;
  int32_t q_cols;
  float q_div;
  int32_t block_n; /* 0 = auto, else 128 or 256 */
  int32_t max_ctas; /* 0 = number of SMs */
  int32_t cta_group; /* 0 = auto, 1 = one CTA per tile, 2 = cta_group::2 pairs (256-row tiles) */
  int32_t k_splits;  /* 0 = auto (fp32-out, EPI_NONE only), 1 = off, n = force n K-splits (atomic fp32 adds) */
  int32_t reserved0_;
  int64_t split_stride; /* > 0 (with k_splits = n > 1): split s STORES its partial sums to C + s*split_stride elements
                           (n deterministic slabs, summed in order by lapb200_resid_norm_fwd) instead of atomic adds */
} lapb_gemm_t;

int lapb200_gemm_bf16(const lapb_gemm_t* p, lapb_stream_t stream);

/* ------------------------------------------------------------------------- */
/* HBM-bound kernels. Convention: every integer is int64_t, every scalar is  */
/* float, bf16 tensors are void*, last argument is the stream.               */
/* ------------------------------------------------------------------------- */

/* fp32 -> bf16 compute copy of parameters (lora.py:57 `w.astype(dtype)` hoisted out of the step). */
int lapb200_cast_f32_bf16(const float* src, void* dst, int64_t n, lapb_stream_t s);
/* dst[v,0:D]=bf16(E[v]), dst[v,D:2D]=bf16(E[v]-hi): the fp32 table of Embedder.decode (gemma.py:153-154) as a hi/lo pair. */
int lapb200_split_hi_lo(const float* src, void* dst, int64_t rows, int64_t D, lapb_stream_t s);
/* images -> fp32 patch rows (siglip.py:216-223 conv as im2col); is_u8 fuses Observation.from_dict's u8/255*2-1 (OP/models/model.py:116-118). */
int lapb200_patchify(const void* img0, const void* img1, const void* img2, int64_t is_u8, float* out, void* out_hi,
                     void* out_lo, int64_t pk_pad, int64_t B, int64_t C, int64_t H, int64_t W, int64_t ps,
                     lapb_stream_t s);
/* fp32 CUDA-core GEMM with generic strides for the layers the reference keeps in fp32:
 * patch conv (siglip.py:216-229, +bias +pos_embedding table), action_in/out_proj, time MLP (pi0.py:159-169, lap.py:298). */
int lapb200_sgemm(const void* A, int64_t a_bf16, const void* B, int64_t b_bf16, void* C, int64_t c_bf16, int64_t M,
                  int64_t N, int64_t K, int64_t sam, int64_t sak, int64_t sbn, int64_t sbk, int64_t ldc,
                  const float* bias, const float* table, int64_t table_rows, int64_t accumulate, lapb_stream_t s);

/* flax nn.LayerNorm(dtype=bf16), eps 1e-6, fp32 fast-variance stats (siglip.py:87,98,161) and its backward
 * (dx = dres + LN'(dy); dscale/dbias accumulated with atomics). */
int lapb200_layernorm_fwd(const void* x, const float* scale, const float* bias, void* y, float* mean, float* rstd,
                          int64_t M, int64_t W, lapb_stream_t s);
int lapb200_layernorm_bwd(const void* dy, const void* x, const float* scale, const float* mean, const float* rstd,
                          const void* dres, void* dx, float* dscale, float* dbias, int64_t M, int64_t W,
                          lapb_stream_t s);

/* gemma.RMSNorm (gemma.py:112-131): plain (scale) or adaptive (mod = [scale|shift|gate] per sample);
 * row_idx gathers source rows, dup writes [y|y] (split-table LM head). */
int lapb200_rmsnorm_fwd(const void* x, int64_t ldx, const int64_t* row_idx, const float* scale, const void* mod,
                        int64_t ldmod, int64_t rows_per_sample, void* y, int64_t ldy, int64_t dup, float* rstd,
                        int64_t M, int64_t D, lapb_stream_t s);
int lapb200_rmsnorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const int64_t* row_idx,
                        const float* scale, const float* rstd, const void* dres, void* dx, float* dscale, int64_t M,
                        int64_t D, lapb_stream_t s);
int lapb200_ada_rmsnorm_bwd(const void* dy, const void* x, const void* mod, int64_t ldmod, const float* rstd,
                            const void* dres, void* dx, void* dmod, int64_t lddmod, int64_t B,
                            int64_t rows_per_sample, int64_t D, lapb_stream_t s);
/* backward of the gated residual x + y*gate (gemma.py:577-583): dy = dxo*gate, dgate = sum_rows dxo*y. */
int lapb200_gated_bwd(const void* dxo, const void* y, const void* gate, int64_t ldg, void* dy, void* dgate,
                      int64_t lddg, int64_t B, int64_t rows_per_sample, int64_t D, lapb_stream_t s);

/* RoPE + q*hd^-0.5 + gather of both experts' fused QKV into attention layout (gemma.py:204,215-218,548-564),
 * also the KV-cache append of the suffix-only pass (gemma.py:227-230); and its backward. */
int lapb200_rope_fwd(const void* qkv0, const void* qkv1, const int32_t* positions, const float* timescale, void* Q,
                     void* Kc, void* Vc, int64_t B, int64_t P, int64_t A, int64_t Tpad, int64_t NH, int64_t HD,
                     int64_t t_begin, float qscale, lapb_stream_t s);
int lapb200_rope_bwd(const void* dQ, const void* dK, const void* dV, const int32_t* positions,
                     const float* timescale, void* dqkv0, void* dqkv1, int64_t B, int64_t P, int64_t A, int64_t Tpad,
                     int64_t NH, int64_t HD, float qscale, lapb_stream_t s);

/* GeGLU backward in place (lora.py:124-142); GELU backward (siglip.py:71); swish (pi0.py:165-167). */
int lapb200_geglu_bwd(void* dact, void* gu, int64_t M, int64_t F, int64_t write_act, lapb_stream_t s);
int lapb200_gelu_bwd(void* dh, const void* pre, int64_t n, lapb_stream_t s);
int lapb200_swish_fwd(const float* z, float* y, void* y_bf16, int64_t n, lapb_stream_t s);
int lapb200_swish_bwd(const float* z, const float* dy, const void* dy_bf16, float* dz, int64_t n, lapb_stream_t s);
/* out[n] += sum_m X[m,n] (bias / pos_embedding gradients). */
int lapb200_colsum(const void* X, int64_t ldx, float* out, int64_t M, int64_t N, lapb_stream_t s);

/* Embedder.encode (gemma.py:148-151,446-448) and its scatter-add backward. */
int lapb200_embed_fwd(const int32_t* ids, const float* E, void* X, int64_t B, int64_t L, int64_t row_off,
                      int64_t rows_per_sample, int64_t D, float scale, lapb_stream_t s);
int lapb200_embed_bwd(const int32_t* ids, const void* dX, float* dE, int64_t B, int64_t L, int64_t row_off,
                      int64_t rows_per_sample, int64_t D, float scale, lapb_stream_t s);
int lapb200_scatter_rows(const void* d, const int64_t* rows, void* dX, int64_t R, int64_t D, lapb_stream_t s);

/* flow-matching inputs x_t, u_t (lap.py:193-197) and posemb_sincos(t) (pi0.py:47-63); Euler step (lap.py:667). */
int lapb200_suffix_inputs(const float* actions, const float* noise, const float* time, float* x_t, float* u_t,
                          float* time_emb, int64_t B, int64_t AD, int64_t W, lapb_stream_t s);
int lapb200_axpy(float* x, const float* v, float dt, int64_t n, lapb_stream_t s);

/* K11: make_attn_mask / combined mask / positions (pi0.py:19-44, lap.py:303-377, :641-654) -> packed bits. */
int lapb200_mask_build(const uint8_t* pm, const uint8_t* par, const uint8_t* pma, const uint8_t* sm,
                       const uint8_t* sar, uint32_t* bits, int32_t* positions, int64_t B, int64_t P, int64_t A,
                       int64_t W32, int64_t row_begin, int64_t infer_rows, lapb_stream_t s);
int lapb200_mask_expand(const uint32_t* bits, uint8_t* dense, int64_t rows, int64_t S, int64_t W32, lapb_stream_t s);
/* masked softmax between the attention GEMMs (gemma.py:258-261) and generic softmax backward. */
int lapb200_attn_softmax_fwd(const float* S, const uint32_t* bits, void* P, int64_t B, int64_t rows_per_batch,
                             int64_t G, int64_t S_len, int64_t ld, int64_t W32, lapb_stream_t s);
int lapb200_softmax_bwd(const void* P, const void* dP, void* dS, int64_t rows, int64_t ld, lapb_stream_t s);
/* SigLIP softmax evaluated in bf16 (flax MultiHeadDotProductAttention, siglip.py:88-93); mode 1 = fp32 softmax. */
int lapb200_vit_softmax_fwd(void* S, int64_t rows, int64_t n, int64_t ld, int64_t mode, lapb_stream_t s);

/* K9: log-softmax cross-entropy over fp32 LM-head logits + weighted bf16 gradient (lap.py:221-260). */
int lapb200_ce_fwd_bwd(const float* logits, int64_t ld, const int32_t* targets, const float* weights, float* nll,
                       void* dlogits, int64_t ldd, int64_t R, int64_t V, lapb_stream_t s);
/* action MSE + gradient (lap.py:291-301); weighted scalar reduction for the final loss (lap.py:573-596). */
int lapb200_mse_fwd_bwd(const float* v, const float* u, float* loss, float* dv, int64_t B, int64_t AD, float gscale,
                        lapb_stream_t s);
int lapb200_weighted_sum(const float* x, const float* w, float* out, int64_t n, float alpha, int64_t accumulate,
                         lapb_stream_t s);

/* K1: fused Gemma shared attention forward (gemma.py:234-272), head_dim 256, GQA 8:1 with the query heads stacked into the
 * MMA M dimension: S = Q K^T and O = P V on tcgen05 with TMEM accumulators, K/V tiles double-buffered in shared
 * memory by TMA, two-pass fp32 softmax with the packed-bit mask (finite -2.3819763e38 fill), P optionally written out
 * (bf16, for the backward).  Rows [0, split_row) of each sample go to O0, the rest to O1 (prefix / action expert). */
int lapb200_fa_gemma_fwd(const void* Q, const void* Kc, const void* Vc, const uint32_t* bits, void* P, void* O0,
                         void* O1, int64_t B, int64_t R, int64_t G, int64_t Tq, int64_t S_len, int64_t Tpad,
                         int64_t W32, int64_t split_row, int64_t head_dim, lapb_stream_t s);

/* Inference (lap.py:634-667): weight-streaming GEMM for M <= 16 rows (the 10 action tokens of one denoise step) with the
 * same fused epilogues as the tile GEMM (NONE/bias, RESID, GATED_RESID, GEGLU), and attention of a few query tokens
 * against the KV cache (gemma.py:227-272). */
int lapb200_skinny_gemm(const void* X, int64_t ldx, const void* W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                        void* Y, int64_t ldy, int64_t y_fp32, int64_t epi, const float* bias, const void* resid,
                        int64_t ldr, const void* gate, int64_t ldg, int64_t gate_rows, void* Y2, int64_t ldy2,
                        lapb_stream_t s);
/* fp32 layers at M <= 16 rows: Y = X W^T + bias (time MLP, action_in/out_proj; pi0.py:159-169, lap.py:665). */
int lapb200_gemv_f32(const void* X, int64_t x_bf16, int64_t ldx, const float* W, const float* bias, void* Y,
                     int64_t ldy, int64_t y_bf16, int64_t M, int64_t N, int64_t K, lapb_stream_t s);
int lapb200_decode_attn(const void* Q, const void* Kc, const void* Vc, const uint32_t* bits, void* O, int64_t B,
                        int64_t Tq, int64_t NH, int64_t HD, int64_t S_len, int64_t Tpad, int64_t W32,
                        lapb_stream_t s);

/* Finalisation of a slab split-K GEMM (lapb_gemm_t.split_stride) fused with the following normalisation, used by the
 * batch-1 prefix pass:  t = bf16(sum_s acc[s]); t = bf16(t + bf16(bias)) if bias; xout = bf16(resid + t) — the rounding
 * points of LAPB_EPI_RESID — then y = LayerNorm(xout) (layernorm = 1: scale, nbias, eps 1e-6; siglip.py:87,98,161) or
 * RMSNorm(xout) = xout*rstd*(1+scale) (gemma.py:112-131); y = NULL: finalisation only.  acc: [nsplit] slabs of [M, D] fp32. */
int lapb200_resid_norm_fwd(const void* resid, const float* acc, int64_t nsplit, int64_t slab_stride, const float* bias,
                           void* xout, int64_t layernorm, const float* scale, const float* nbias, void* y, float* mean,
                           float* rstd, int64_t M, int64_t D, lapb_stream_t s);

/* delta[(bo*nbi + bi)*out_rows + out_off + r] = sum_d dO[bo, bi, r, d] * O[bo, bi, r, d] (bf16 in, fp32 out): the row
 * term rowsum(P o dP) = dO . O of the softmax backward, consumed by LAPB_EPI_SOFTMAX_BWD.  Rows are `ldd` / `ldo` elements
 * apart, the two batch levels `*_bs_i` / `*_bs_o` elements apart. */
int lapb200_rowdot(const void* dO, const void* O, float* delta, int64_t rows, int64_t D, int64_t ldd, int64_t ldo,
                   int64_t nbi, int64_t nbo, int64_t d_bs_i, int64_t d_bs_o, int64_t o_bs_i, int64_t o_bs_o, int64_t out_rows,
                   int64_t out_off, lapb_stream_t s);

/* Image side of preprocess_observation (src/lap/models/model_adapter.py:83-181), ahead of lapb200_patchify.
 * image_resize_pad: resize_with_pad (model_adapter.py:113-116 -> OP/shared/image_tools.py:11-52) of uint8 or float32
 *   [B, Hin, Win, 3] images into float32 [B, Hout, Wout, 3] in [-1, 1] (uint8: round, clip, then the u8/255*2-1 of
 *   Observation.from_dict).  The separable antialiasing filter of jax.image.resize is passed as sparse rows: output row p of
 *   the resized region [rh, rw] (placed at (ph0, pw0)) reads input rows ystart[p] .. +ytaps-1 with weights yw[p*ytaps + t].
 * image_augment: model_adapter.py:118-151 with explicit per-sample parameters params[b] = (crop_y, crop_x, angle_deg,
 *   brightness, contrast, saturation, skip, unused) — one bilinear resampling (crop 95 % -> resize -> rotate) + colour jitter;
 *   definition: oracle/image_oracle.py.  src uint8 (0..255) or float32 in [-1, 1]; dst float32 in [-1, 1], dst != src. */
int lapb200_image_resize_pad(const void* src, int64_t src_is_u8, float* dst, int64_t B, int64_t Hin, int64_t Win,
                             int64_t Hout, int64_t Wout, int64_t rh, int64_t rw, int64_t ph0, int64_t pw0,
                             const int32_t* ystart, const float* yw, int64_t ytaps, const int32_t* xstart, const float* xw,
                             int64_t xtaps, lapb_stream_t s);
int lapb200_image_augment(const void* src, int64_t src_is_u8, float* dst, int64_t B, int64_t H, int64_t W,
                          const float* params, lapb_stream_t s);

/* K2: fused SigLIP attention forward (flax MultiHeadDotProductAttention between the QKV and output projections,
 * OP/models/siglip.py:88-93) for head_dim in (64, 80] (So400m: 72).  qkv: bf16 [Ni*Np, 3, nh, hd] rows = (image, token), q
 * already divided by sqrt(hd); O: bf16 [Ni*Np, nh*hd]; P (optional, Np % 8 == 0): bf16 [Ni, nh, Np, Np] softmax
 * probabilities saved for the backward pass.  mode 0: softmax evaluated in bf16 like flax 0.10.2 (bf16 logits, e =
 * bf16(exp(bf16(s - max))), bf16(sum), p = bf16(e / sum)); mode 1: fp32 softmax of the bf16 logits, one rounding of P. */
int lapb200_vit_attn_fwd(const void* qkv, void* O, void* P, int64_t Ni, int64_t nh, int64_t Np, int64_t hd, int64_t mode,
                         lapb_stream_t s);

/* K10: the whole flow-matching Euler loop of LAP.sample_actions (lap.py:634-672: embed_suffix pi0.py:139-186, the
 * action-expert half of gemma.Module gemma.py:455-531 against the prefix KV cache, action_out_proj, x += dt*v) for ONE
 * sample as ONE persistent cooperative kernel (one CTA per SM, grid barriers between dependent phases, next-phase weights
 * prefetched into L2 while waiting).  All pointers are device pointers; bf16 tensors are row-major [out, in] weights /
 * [rows, features] activations; `*_ls` = element stride between consecutive layers.  Scratch buffers are caller-owned. */
#define LAPB_DENOISE_SYNC_WORDS 32   /* uint32 words behind `sync` */
typedef struct {
  int32_t A, ad, D1, NH, HD, F1, L, Pn, Tpad, TpadK, W32, nm, num_steps; /* TpadK = round_up(Pn, 64) = row length of VcT */
  float dt, qscale;                  /* Euler step (-1/num_steps), head_dim^-0.5 */
  float times[16];                   /* t of every step (1, 1+dt, ...) */
  const void *qkv_w, *o_w, *gu_w, *down_w; /* expert weights, layer 0: [(NH+2)HD, D1] [D1, NH*HD] [2*F1, D1] [D1, F1] */
  int64_t qkv_ls, o_ls, gu_ls, down_ls;
  const void* mod_w;                 /* [nm*3*D1, D1] bf16: all adaRMS modulation Dense layers stacked */
  const float* mod_b;                /* [nm*3*D1] */
  const float *ain_w, *ain_b, *tin_w, *tin_b, *tout_w, *tout_b, *aout_w, *aout_b; /* fp32 suffix projections */
  const void* Kc;                    /* [L][Tpad][HD] bf16 prefix keys (post-RoPE) */
  const void* VcT;                   /* [L][HD][TpadK] bf16 prefix values, transposed (lapb200_transpose_v) */
  int64_t kc_ls, vct_ls;
  const uint32_t* bits;              /* [A][W32] packed attention mask of the suffix rows over P+A keys */
  const int32_t* pos;                /* [A] RoPE positions of the suffix rows */
  const float* timescale;            /* [HD/2] */
  float* x;                          /* [A*ad] in: noise, out: actions */
  float* s1;                         /* scratch [num_steps*D1] */
  void *cond16, *mod;                /* scratch bf16 [num_steps*D1], [num_steps * nm*3*D1] */
  void *XE, *XE1, *qkv, *O, *act;    /* scratch bf16 [16*D1] x2, [16*(NH+2)*HD], [16*NH*HD], [16*F1] */
  float *part_o, *part_ml;           /* scratch [NH*(TpadK/64+1)*16*HD], [NH*(TpadK/64+1)*16*2] */
  uint32_t* sync;                    /* [LAPB_DENOISE_SYNC_WORDS], zeroed by the launcher: [0] barrier counter, [1] error flag
                                        (set if a barrier timed out), [2..] per-head arrival counters (LAPB_DENOISE_FOLD=1) */
  unsigned long long* prof;          /* optional [32]: ns per phase slot seen by CTA 0 (16.. = sub-phase marks); (0 prologue, 1 action_in, 2/3 P1 work /
                                        barrier, 4/5 P2, 6/7 P2b, 8/9 P3, 10/11 P4, 12/13 P5, 14 final); NULL = off */
  int32_t packed;                    /* 1: qkv_w / o_w / gu_w / down_w, Kc ([Tpad, HD] per layer) and VcT ([HD, TpadK] per layer)
                                        are TILE-MAJOR copies made by lapb200_pack_tiles (cluster kernel only); 0: row-major */
  int32_t flags;                     /* experiments / tests, 0 = product path.  v2 loop kernel: bit 1 = fence.sc + ld.acquire
                                        grid barrier.  bit 7 = round-1 kernel layout (also the fallback when v2's shared
                                        memory does not fit), which reads bit 0 = layer-ahead L2 prefetch burst, bit 1 =
                                        fence after the register preloads, bit 2 = per-phase L2 prefetch */
} lapb_denoise_params_t;
/* 1 if the shape is supported by the persistent kernel (B == 1, A <= 16, num_steps <= 16, head_dim <= 256 ...). */
int lapb200_denoise_supported(int64_t B, int64_t A, int64_t ad, int64_t D1, int64_t NH, int64_t HD, int64_t F1,
                              int64_t Pn, int64_t Tpad, int64_t num_steps);
int lapb200_denoise_grid(void);
int lapb200_denoise_loop(const lapb_denoise_params_t* params, lapb_stream_t s);
/* Tile-major copy of a [rows, cols] bf16 matrix for the weight-streaming kernels (rows % 8 == 0, cols % 32 == 0):
 * dst[((r/8 * cols/32 + c/32) * 8 + r%8) * 32 + c%32] = src[r*row_stride + c*col_stride] (0 for c >= valid_cols), `batch`
 * matrices `src_bs` / `dst_bs` elements apart.  A warp of the streaming kernels then reads 512 contiguous bytes per load. */
int lapb200_pack_tiles(const void* src, void* dst, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride,
                       int64_t valid_cols, int64_t batch, int64_t src_bs, int64_t dst_bs, lapb_stream_t s);
/* VcT[l][d][j] = Vc[l][j][d] for j < Pn, 0 for Pn <= j < TpadK. */
int lapb200_transpose_v(const void* Vc, void* VcT, int64_t L, int64_t Tpad, int64_t TpadK, int64_t HD, int64_t Pn,
                        lapb_stream_t s);

/* buf[dst] = sqrt(buf[src]) on the device (param_norm = sqrt(sum p^2), scripts/train.py:411) */
int lapb200_sqrt_scalar(float* buf, int64_t src, int64_t dst, lapb_stream_t s);

/* K12: global-norm clip + AdamW + EMA + bf16 copy + norms over the flat state
 * (scripts/train.py:363-415; OP/training/optimizer.py:76-85). */
int lapb200_opt_num_partials(void);
int lapb200_sumsq_partials(const float* x, int64_t n, float* partials, lapb_stream_t s);
int lapb200_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, void* w16, int64_t n,
                      const float* gpartials, int64_t n_partials, float* stats, int64_t kernel_begin,
                      int64_t kernel_end, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2,
                      float clip, float ema_decay, int64_t ema_on, const float* hyper, lapb_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* LAPB200_H_ */
