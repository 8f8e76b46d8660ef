// Inference-side kernels for the flow-matching denoise loop (lap.py:634-667), where the action expert sees only
// M = batch * action_horizon = 10 rows per step: every projection is a weight-streaming, HBM-bound "skinny" GEMM and
// the attention is 80 query rows against a ~700-key cache.
//
//   skinny_gemm  : Y[M<=16, N] = epi( X[M,K] W[N,K]^T ).  The tcgen05 tile kernel would run 4-10 CTAs here; this
//                  kernel instead spreads the weight rows over N/8 CTAs x 8 warps (the warps split K), loads weights
//                  with 128-bit streaming loads and multiplies with mma.sync m16n8k16 (bf16 in, fp32 accumulate) —
//                  the tensor pipe is irrelevant at M=10, the roofline is HBM bandwidth (bytes of W).
//                  Fused epilogues: bias, residual, gated residual (+branch output), GeGLU (gate/up pairs).
//   decode_attn  : softmax(mask(q K^T)) V for a handful of query tokens against the KV cache, one CTA per
//                  (sample, query token, head); fp32 logits/softmax like gemma.py:235-271.
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"

namespace lapb {

typedef __nv_bfloat16 bf16;
#define BIG_NEG (-2.3819763e38f)

__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {  // weights are read exactly once: do not pollute L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct SkinnyArgs {
  const bf16* X; long ldx;
  const bf16* W; long ldw;
  int M, N, K;
  void* Y; long ldy; int y_fp32;
  int epi;
  const float* bias;
  const bf16* resid; long ldr;
  const bf16* gate; long ldg; int gate_rows;
  bf16* Y2; long ldy2;
};

// NT = n8 tiles per CTA.  GeGLU uses NT=2: tile 0 = gate columns n0.., tile 1 = up columns N + n0.. of the stacked weight.
template <int NT>
__global__ void __launch_bounds__(256) skinny_gemm_kernel(SkinnyArgs a) {
  __shared__ float red[8][NT][32][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const bool dual = a.epi == LAPB_EPI_GEGLU;
  const int n0 = blockIdx.x * (dual ? 8 : 8 * NT);
  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bf16* xlo = a.X + (long)g * a.ldx + 8 * t;
  const bf16* xhi = a.X + (long)(g + 8) * a.ldx + 8 * t;
  const bool vlo = g < a.M, vhi = (g + 8) < a.M;
  const bf16* wrow[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    int col = dual ? (n0 + i * a.N) : (n0 + 8 * i);
    wrow[i] = a.W + (long)(col + g) * a.ldw + 8 * t;
  }
  const int ngroups = a.K >> 5;  // 32 K-elements per group; warps interleave groups
  constexpr int UNROLL = 4;
  for (int kg0 = warp; kg0 < ngroups; kg0 += 8 * UNROLL) {
    uint4 b[UNROLL][NT], alo[UNROLL], ahi[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      int kg = kg0 + 8 * u;
      if (kg < ngroups) {
#pragma unroll
        for (int i = 0; i < NT; ++i) b[u][i] = ld_stream_u4(wrow[i] + kg * 32);
        alo[u] = vlo ? *reinterpret_cast<const uint4*>(xlo + kg * 32) : make_uint4(0, 0, 0, 0);
        ahi[u] = vhi ? *reinterpret_cast<const uint4*>(xhi + kg * 32) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      int kg = kg0 + 8 * u;
      if (kg < ngroups) {
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          // K is consumed in a permuted order (lane t owns elements 8t..8t+7 of the group) — the same permutation
          // on both operands, so the dot products are unchanged and every load is a contiguous 16 bytes.
          mma_bf16_16816(acc[i], alo[u].x, ahi[u].x, alo[u].y, ahi[u].y, b[u][i].x, b[u][i].y);
          mma_bf16_16816(acc[i], alo[u].z, ahi[u].z, alo[u].w, ahi[u].w, b[u][i].z, b[u][i].w);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[warp][i][lane][j] = acc[i][j];
  __syncthreads();
  // one thread per output element of the [16 x 8*NT] tile
  const int ncols = 8 * NT;
  const int e = threadIdx.x;
  if (e >= 16 * ncols) return;
  const int m = e / ncols, c = e % ncols;
  if (m >= a.M) return;
  auto tile_val = [&](int tile, int cc) {
    int src_lane = (m & 7) * 4 + (cc >> 1), idx = (m >> 3) * 2 + (cc & 1);
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][tile][src_lane][idx];
    return s;
  };
  if (dual) {
    if (c >= 8) return;
    const int n = n0 + c;
    if (n >= a.N) return;
    float gv = bf16r(tile_val(0, c)), uv = bf16r(tile_val(1, c));
    reinterpret_cast<bf16*>(a.Y)[(long)m * a.ldy + n] = __float2bfloat16_rn(bf16r(gelu_tanh(gv)) * uv);
    if (a.Y2) {
      a.Y2[(long)m * a.ldy2 + n] = __float2bfloat16_rn(gv);
      a.Y2[(long)m * a.ldy2 + a.N + n] = __float2bfloat16_rn(uv);
    }
    return;
  }
  const int n = n0 + c;
  if (n >= a.N) return;
  float v = tile_val(c >> 3, c & 7);
  if (a.y_fp32) {
    if (a.bias) v += a.bias[n];
    reinterpret_cast<float*>(a.Y)[(long)m * a.ldy + n] = v;
    return;
  }
  if (a.bias) v = bf16r(v) + bf16r(a.bias[n]);
  if (a.epi == LAPB_EPI_RESID) {
    v = __bfloat162float(a.resid[(long)m * a.ldr + n]) + bf16r(v);
  } else if (a.epi == LAPB_EPI_GATED_RESID) {
    float y = bf16r(v);
    if (a.Y2) a.Y2[(long)m * a.ldy2 + n] = __float2bfloat16_rn(y);
    float gt = __bfloat162float(a.gate[(long)(m / a.gate_rows) * a.ldg + n]);
    v = __bfloat162float(a.resid[(long)m * a.ldr + n]) + bf16r(y * gt);
  }
  reinterpret_cast<bf16*>(a.Y)[(long)m * a.ldy + n] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------
// decode attention: Q [B, Tq, NH, HD] (already RoPE'd and scaled), K/V cache [B, Tpad, HD] (one KV head),
// mask bits [B, Tq, W32], O [B, Tq, NH, HD].  One CTA (128 threads) per (b, tq, head).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
decode_attn_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ Kc, const bf16* __restrict__ Vc,
                   const uint32_t* __restrict__ bits, bf16* __restrict__ O, int Tq, int NH, int HD, int S_len,
                   int Tpad, int W32) {
  extern __shared__ float sh[];  // [HD] q | [Tpad] probs | [32] red | [8][HD] partial outputs
  float* qs = sh;
  float* ps = sh + HD;
  float* red = ps + Tpad;
  float* part = red + 32;
  const int h = blockIdx.x % NH;
  const int tq = (blockIdx.x / NH) % Tq;
  const int b = blockIdx.x / (NH * Tq);
  const bf16* q = Q + (((long)b * Tq + tq) * NH + h) * HD;
  const bf16* K = Kc + (long)b * Tpad * HD;
  const bf16* V = Vc + (long)b * Tpad * HD;
  const uint32_t* mrow = bits + ((long)b * Tq + tq) * W32;
  for (int d = threadIdx.x; d < HD; d += 256) qs[d] = __bfloat162float(q[d]);
  __syncthreads();
  // ---- logits: one warp per key (coalesced 16-byte loads across the head dimension), 8 keys in flight per warp ----
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float lmax = -3.4e38f;
  for (int j0 = warp * 8; j0 < S_len; j0 += 64) {
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc[u] = 0.f;
      int j = j0 + u;
      if (j < S_len) {
        for (int d = lane * 8; d < HD; d += 256) {
          uint4 kv = *reinterpret_cast<const uint4*>(K + (long)j * HD + d);
          float2 f;
          f = unpack_bf16x2(kv.x); acc[u] += f.x * qs[d] + f.y * qs[d + 1];
          f = unpack_bf16x2(kv.y); acc[u] += f.x * qs[d + 2] + f.y * qs[d + 3];
          f = unpack_bf16x2(kv.z); acc[u] += f.x * qs[d + 4] + f.y * qs[d + 5];
          f = unpack_bf16x2(kv.w); acc[u] += f.x * qs[d + 6] + f.y * qs[d + 7];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = warp_sum(acc[u]);
    if (lane < 8) {
      int j = j0 + lane;
      if (j < S_len) {
        float sc = acc[0];
#pragma unroll
        for (int u = 1; u < 8; ++u) sc = (lane == u) ? acc[u] : sc;
        bool ok = (mrow[j >> 5] >> (j & 31)) & 1u;
        sc = ok ? sc : BIG_NEG;
        ps[j] = sc;
        lmax = fmaxf(lmax, sc);
      }
    }
  }
  __syncthreads();
  float m = block_max(lmax, red);
  float lsum = 0.f;
  for (int j = threadIdx.x; j < S_len; j += 256) {
    float e = __expf(ps[j] - m);
    ps[j] = e;
    lsum += e;
  }
  float sum = block_sum(lsum, red);
  float inv = 1.0f / sum;
  __syncthreads();
  // probabilities are rounded to bf16 before the PV product (gemma.py:261)
  for (int j = threadIdx.x; j < S_len; j += 256) ps[j] = bf16r(ps[j] * inv);
  __syncthreads();
  // ---- O = P V: thread = (8-dim group dg, key group kg); 8 independent V-row loads in flight per thread ----
  const int ndg = HD / 8;             // dim groups (32 for HD = 256)
  const int nkg = 256 / ndg;          // key groups sharing the CTA (8 for HD = 256)
  const int dg = threadIdx.x % ndg, kg = threadIdx.x / ndg;
  float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (kg < nkg) {
    for (int j0 = kg; j0 < S_len; j0 += nkg * 8) {
      uint4 vv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        int j = j0 + u * nkg;
        vv[u] = (j < S_len) ? *reinterpret_cast<const uint4*>(V + (long)j * HD + dg * 8) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        int j = j0 + u * nkg;
        float p = (j < S_len) ? ps[j] : 0.f;
        float2 f;
        f = unpack_bf16x2(vv[u].x); o[0] += p * f.x; o[1] += p * f.y;
        f = unpack_bf16x2(vv[u].y); o[2] += p * f.x; o[3] += p * f.y;
        f = unpack_bf16x2(vv[u].z); o[4] += p * f.x; o[5] += p * f.y;
        f = unpack_bf16x2(vv[u].w); o[6] += p * f.x; o[7] += p * f.y;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) part[kg * HD + dg * 8 + i] = o[i];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < HD; d += 256) {
    float acc = 0.f;
    for (int k2 = 0; k2 < nkg; ++k2) acc += part[k2 * HD + d];
    O[(((long)b * Tq + tq) * NH + h) * HD + d] = __float2bfloat16_rn(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 GEMV-like kernel for the reference's fp32 layers at M <= 16 rows (time MLP, action_in/out_proj; pi0.py:159-169,
// lap.py:665): Y[M, N] = X[M, K] W[N, K]^T + bias.  One warp per output column; fp32 weights streamed once.
// ------------------------------------------------------------------------------------------------
template <typename TX>
__global__ void __launch_bounds__(256)
gemv_f32_kernel(const TX* __restrict__ X, long ldx, const float* __restrict__ W, const float* __restrict__ bias,
                void* __restrict__ Y, long ldy, int y_bf16, int M, int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  float acc[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) acc[m] = 0.f;
  const float* w = W + (long)n * K;
  for (int k = lane; k < K; k += 32) {
    float wv = w[k];
#pragma unroll
    for (int m = 0; m < 16; ++m)
      if (m < M) acc[m] += wv * (float)X[(long)m * ldx + k];
  }
#pragma unroll
  for (int m = 0; m < 16; ++m) acc[m] = warp_sum(acc[m]);
  if (lane == 0) {
    float b = bias ? bias[n] : 0.f;
    for (int m = 0; m < M; ++m) {
      float v = acc[m] + b;
      if (y_bf16) reinterpret_cast<bf16*>(Y)[(long)m * ldy + n] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(Y)[(long)m * ldy + n] = v;
    }
  }
}

}  // namespace lapb

using namespace lapb;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int lapb200_skinny_gemm(const void* X, int64_t ldx, const void* W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                        void* Y, int64_t ldy, int64_t y_fp32, int64_t epi, const float* bias, const void* resid,
                        int64_t ldr, const void* gate, int64_t ldg, int64_t gate_rows, void* Y2, int64_t ldy2,
                        lapb_stream_t s) {
  LAPB_REQUIRE(M >= 1 && M <= 16, "skinny_gemm: M must be in [1,16] (got %ld)", (long)M);
  LAPB_REQUIRE(K % 32 == 0 && N % 8 == 0, "skinny_gemm: K %% 32 and N %% 8 must be 0 (K=%ld N=%ld)", (long)K, (long)N);
  LAPB_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0, "skinny_gemm: ldx, ldw must be multiples of 8");
  LAPB_REQUIRE(epi == LAPB_EPI_NONE || epi == LAPB_EPI_RESID || epi == LAPB_EPI_GATED_RESID || epi == LAPB_EPI_GEGLU,
               "skinny_gemm: unsupported epilogue %ld", (long)epi);
  SkinnyArgs a;
  a.X = (const bf16*)X; a.ldx = ldx; a.W = (const bf16*)W; a.ldw = ldw;
  a.M = (int)M; a.N = (int)N; a.K = (int)K;
  a.Y = Y; a.ldy = ldy; a.y_fp32 = (int)y_fp32; a.epi = (int)epi; a.bias = bias;
  a.resid = (const bf16*)resid; a.ldr = ldr; a.gate = (const bf16*)gate; a.ldg = ldg;
  a.gate_rows = gate_rows > 0 ? (int)gate_rows : 1;
  a.Y2 = (bf16*)Y2; a.ldy2 = ldy2;
  if (epi == LAPB_EPI_GEGLU) {
    skinny_gemm_kernel<2><<<(unsigned)(N / 8), 256, 0, STREAM(s)>>>(a);
  } else {
    skinny_gemm_kernel<1><<<(unsigned)(N / 8), 256, 0, STREAM(s)>>>(a);
  }
  LAPB_LAUNCH_OK("skinny_gemm");
  return 0;
}

int lapb200_gemv_f32(const void* X, int64_t x_bf16, int64_t ldx, const float* W, const float* bias, void* Y,
                     int64_t ldy, int64_t y_bf16, int64_t M, int64_t N, int64_t K, lapb_stream_t s) {
  LAPB_REQUIRE(M >= 1 && M <= 16, "gemv_f32: M must be in [1,16]");
  if (x_bf16)
    gemv_f32_kernel<bf16><<<cdiv(N, 8), 256, 0, STREAM(s)>>>((const bf16*)X, ldx, W, bias, Y, ldy, (int)y_bf16,
                                                             (int)M, (int)N, (int)K);
  else
    gemv_f32_kernel<float><<<cdiv(N, 8), 256, 0, STREAM(s)>>>((const float*)X, ldx, W, bias, Y, ldy, (int)y_bf16,
                                                              (int)M, (int)N, (int)K);
  LAPB_LAUNCH_OK("gemv_f32");
  return 0;
}

int lapb200_decode_attn(const void* Q, const void* Kc, const void* Vc, const uint32_t* bits, void* O, int64_t B,
                        int64_t Tq, int64_t NH, int64_t HD, int64_t S_len, int64_t Tpad, int64_t W32,
                        lapb_stream_t s) {
  LAPB_REQUIRE(HD % 8 == 0 && HD <= 2048 && 256 % (HD / 8) == 0 && S_len <= Tpad, "decode_attn: bad head_dim / lengths");
  size_t smem = (size_t)(HD + Tpad + 32 + (256 / (HD / 8)) * HD) * sizeof(float);
  decode_attn_kernel<<<(unsigned)(B * Tq * NH), 256, smem, STREAM(s)>>>((const bf16*)Q, (const bf16*)Kc,
                                                                         (const bf16*)Vc, bits, (bf16*)O, (int)Tq,
                                                                         (int)NH, (int)HD, (int)S_len, (int)Tpad,
                                                                         (int)W32);
  LAPB_LAUNCH_OK("decode_attn");
  return 0;
}

}  // extern "C"
