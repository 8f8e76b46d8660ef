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
  LAPB_EPI_QSCALE = 5         /* as NONE(+bias); columns < q_cols are divided by q_div      flax MHA q/sqrt(d) */
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
} lapb_gemm_t;

int lapb200_gemm_bf16(const lapb_gemm_t* p, lapb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LAPB200_H_ */
