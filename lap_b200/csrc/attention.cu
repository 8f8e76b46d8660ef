// K11 + the softmax halves of K1/K2: attention-mask / position construction (bit-exact integer work) and the
// masked-softmax forward/backward that sit between the two tcgen05 attention GEMMs (S = Q K^T, O = P V).
//
//   mask_build      OP/models/pi0.py:19-44 (make_attn_mask), lap.py:303-377 (prefix/action masks, positions),
//                   lap.py:641-654 (inference suffix rows)
//   attn_softmax    gemma.py:258-261: where(mask, logits, -2.3819763e38) -> softmax fp32 -> bf16
//   vit_softmax     flax MultiHeadDotProductAttention softmax evaluated in bf16 (siglip.py:88-93)
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"

namespace lapb {

typedef __nv_bfloat16 bf16;
#define BIG_NEG (-2.3819763e38f)

// ------------------------------------------------------------------------------------------------
// mask + positions. One CTA per sample.
//   prefix rows i<P:   allowed(i,j) = j<P & pm[i] & pm[j] & cumP[j] <= cumP[i]
//   suffix rows i>=P:  input = [pma | sm], ar = [0.. | sar];  allowed = in[i] & in[j] & cum[j] <= cum[i]
//                      (infer_rows: prefix part is pm[j] alone, lap.py:644)
//   positions: prefix cumsum(pm)-1 ; suffix sum(pma) + cumsum(sm) - 1
// bits[b, i - row_begin, w] bit k  <->  key j = 32w + k
// ------------------------------------------------------------------------------------------------
// block-wide inclusive scan of v[0..n) (n <= 4 * 256) in shared memory: 4 consecutive elements per thread, warp shuffles
__device__ __forceinline__ void block_scan_inclusive(int* v, int n, int* warp_tot /*[8]*/) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  for (int base = 0; base < n; base += 1024) {  // chunks of 1024, carried through warp_tot[8]
    int x[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = base + 4 * t + k;
      x[k] = (j < n) ? v[j] : 0;
      sum += x[k];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    __syncthreads();
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int off = (base > 0) ? warp_tot[8] : 0;
    for (int k = 0; k < w; ++k) off += warp_tot[k];
    int run = off + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = base + 4 * t + k;
      run += x[k];
      if (j < n) v[j] = run;
    }
    __syncthreads();
    if (t == 255) warp_tot[8] = run;  // carry into the next chunk
    __syncthreads();
  }
}

// grid = (B, row chunks): every CTA rebuilds the (cheap) scans of its sample and writes the mask words of its rows.
__global__ void __launch_bounds__(256)
mask_build_kernel(const uint8_t* __restrict__ pm, const uint8_t* __restrict__ par, const uint8_t* __restrict__ pma,
                  const uint8_t* __restrict__ sm, const uint8_t* __restrict__ sar, uint32_t* __restrict__ bits,
                  int* __restrict__ positions, int P, int A, int W32, int row_begin, int infer_rows) {
  extern __shared__ int sh[];
  __shared__ int warp_tot[9];
  __shared__ int na_s;
  int T = P + A;
  int* cum = sh;            // [T]  cumsum of ar for prefix rows' view (prefix part) / suffix view (suffix part)
  int* valid_p = sh + T;    // [T]  pm (prefix view); suffix part = 0
  int* valid_a = sh + 2 * T;  // [T] pma | sm (action view)
  int* pos = sh + 3 * T;    // [T]
  int b = blockIdx.x;
  const uint8_t* pmb = pm + (long)b * P;
  const uint8_t* parb = par + (long)b * P;
  const uint8_t* pmab = pma ? pma + (long)b * P : pmb;
  for (int j = threadIdx.x; j < P; j += 256) {
    cum[j] = parb[j] ? 1 : 0;
    valid_p[j] = pmb[j] ? 1 : 0;
    valid_a[j] = pmab[j] ? 1 : 0;
    pos[j] = valid_p[j];
  }
  __syncthreads();
  block_scan_inclusive(cum, P, warp_tot);   // cumsum(ar) over the prefix
  block_scan_inclusive(pos, P, warp_tot);   // cumsum(pm)
  {  // na = sum(pma): block reduction
    int part = 0;
    for (int j = threadIdx.x; j < P; j += 256) part += valid_a[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      int na = 0;
      for (int k = 0; k < 8; ++k) na += warp_tot[k];
      na_s = na;
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < P; j += 256) pos[j] -= 1;
  if (threadIdx.x == 0) {  // the suffix is a handful of tokens
    int cs = 0, sc = 0;
    for (int j = 0; j < A; ++j) {
      cs += sar[(long)b * A + j] ? 1 : 0;
      cum[P + j] = cs;
      valid_p[P + j] = 0;
      valid_a[P + j] = sm[(long)b * A + j] ? 1 : 0;
      sc += valid_a[P + j];
      pos[P + j] = na_s + sc - 1;
    }
  }
  __syncthreads();
  int nrows = T - row_begin;
  if (blockIdx.y == 0)
    for (int i = row_begin + threadIdx.x; i < T; i += 256) positions[(long)b * nrows + (i - row_begin)] = pos[i];
  // rows of this CTA: a contiguous slice; one thread per 32-key word
  const int rows_per = (nrows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(nrows, r0 + rows_per);
  long total = (long)max(0, r1 - r0) * W32;
  for (long idx = threadIdx.x; idx < total; idx += 256) {
    int i = row_begin + r0 + (int)(idx / W32), w = (int)(idx % W32);
    uint32_t word = 0;
    for (int k = 0; k < 32; ++k) {
      int j = w * 32 + k;
      if (j >= T) break;
      bool ok;
      if (i < P) {
        ok = (j < P) && valid_p[i] && valid_p[j] && (cum[j] <= cum[i]);
      } else {
        int cj = (j < P) ? 0 : cum[j];
        if (infer_rows && j < P)
          ok = valid_p[j] != 0;
        else
          ok = valid_a[i] && valid_a[j] && (cj <= cum[i]);
      }
      word |= (ok ? 1u : 0u) << k;
    }
    bits[((long)b * nrows + (i - row_begin)) * W32 + w] = word;
  }
}

// expand packed bits to a dense uint8 mask (tests / debugging): dense[b,i,j]
__global__ void mask_expand_kernel(const uint32_t* __restrict__ bits, uint8_t* __restrict__ dense, long rows, int S,
                                   int W32) {
  long total = rows * S;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    long r = idx / S;
    int j = idx % S;
    dense[idx] = (bits[r * W32 + (j >> 5)] >> (j & 31)) & 1u;
  }
}

// ------------------------------------------------------------------------------------------------
// Gemma masked softmax. S fp32 [B, R, ld] (R = Tq*G rows, G query heads share a mask row), one warp per row.
// P bf16 [B, R, ld] gets zeros in the pad columns [S_len, ld).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_softmax_fwd_kernel(const float* __restrict__ S, const uint32_t* __restrict__ bits, bf16* __restrict__ Pout,
                        long rows_per_batch, int G, int S_len, int ld, int W32, long total_rows) {
  long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= total_rows) return;
  int lane = threadIdx.x & 31;
  long b = row / rows_per_batch, r = row % rows_per_batch;
  const uint32_t* mrow = bits + (b * (rows_per_batch / G) + r / G) * W32;
  const float4* s4 = reinterpret_cast<const float4*>(S + row * ld);
  const int ngroups = ld >> 2;  // float4 groups per row (ld % 32 == 0)
  float4 v[8];
  float mx = -3.4e38f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int g = lane + 32 * i;
    if (g < ngroups) {
      float4 x = s4[g];
      int c = g * 4;
      uint32_t m4 = (mrow[c >> 5] >> (c & 31)) & 0xFu;
      x.x = (m4 & 1u) ? x.x : BIG_NEG;
      x.y = (m4 & 2u) ? x.y : BIG_NEG;
      x.z = (m4 & 4u) ? x.z : BIG_NEG;
      x.w = (m4 & 8u) ? x.w : BIG_NEG;
      // columns >= S_len are padding: excluded from max / sum, written as 0
      if (c + 0 < S_len) mx = fmaxf(mx, x.x);
      if (c + 1 < S_len) mx = fmaxf(mx, x.y);
      if (c + 2 < S_len) mx = fmaxf(mx, x.z);
      if (c + 3 < S_len) mx = fmaxf(mx, x.w);
      v[i] = x;
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int g = lane + 32 * i;
    if (g < ngroups) {
      int c = g * 4;
      float4 e;
      e.x = (c + 0 < S_len) ? __expf(v[i].x - mx) : 0.f;
      e.y = (c + 1 < S_len) ? __expf(v[i].y - mx) : 0.f;
      e.z = (c + 2 < S_len) ? __expf(v[i].z - mx) : 0.f;
      e.w = (c + 3 < S_len) ? __expf(v[i].w - mx) : 0.f;
      sum += (e.x + e.y) + (e.z + e.w);
      v[i] = e;
    }
  }
  sum = warp_sum(sum);
  float inv = 1.0f / sum;
  uint2* p2 = reinterpret_cast<uint2*>(Pout + row * ld);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int g = lane + 32 * i;
    if (g < ngroups) {
      uint2 o;
      o.x = pack_bf16x2(v[i].x * inv, v[i].y * inv);
      o.y = pack_bf16x2(v[i].z * inv, v[i].w * inv);
      p2[g] = o;
    }
  }
}

// dS = P * (dP - sum_j P_j dP_j);  P, dP, dS bf16 [rows, ld]; one warp per row (in place on dP allowed)
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const bf16* __restrict__ P, const bf16* __restrict__ dP, bf16* __restrict__ dS, int ld,
                   long total_rows) {
  long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= total_rows) return;
  int lane = threadIdx.x & 31;
  const bf16* p = P + row * ld;
  const bf16* d = dP + row * ld;
  float dot = 0.f;
  for (int c = lane * 2; c < ld; c += 64) {
    float2 pv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p + c));
    float2 dv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(d + c));
    dot += pv.x * dv.x + pv.y * dv.y;
  }
  dot = warp_sum(dot);
  for (int c = lane * 2; c < ld; c += 64) {
    float2 pv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p + c));
    float2 dv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(d + c));
    *reinterpret_cast<uint32_t*>(dS + row * ld + c) = pack_bf16x2(pv.x * (dv.x - dot), pv.y * (dv.y - dot));
  }
}

// ------------------------------------------------------------------------------------------------
// SigLIP softmax in bf16 arithmetic (in place), one warp per row of `n` columns (n % 2 == 0):
//   e = bf16(exp(bf16(x - max)));  p = bf16(e / bf16(sum e))          [mode 0: as written in jax.nn.softmax on bf16]
//   p = bf16(softmax_fp32(x))                                          [mode 1]
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
vit_softmax_fwd_kernel(bf16* __restrict__ S, int n, int ld, long total_rows, int mode) {
  long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= total_rows) return;
  int lane = threadIdx.x & 31;
  bf16* s = S + row * ld;
  const int nvec = n >> 3;  // 8 bf16 per 16-byte vector (n % 8 == 0)
  float x[NV][8];
  float mx = -3.4e38f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int v = lane + 32 * i;
    if (v < nvec) {
      uint4 u = *reinterpret_cast<const uint4*>(s + v * 8);
      float2 f;
      f = unpack_bf16x2(u.x); x[i][0] = f.x; x[i][1] = f.y;
      f = unpack_bf16x2(u.y); x[i][2] = f.x; x[i][3] = f.y;
      f = unpack_bf16x2(u.z); x[i][4] = f.x; x[i][5] = f.y;
      f = unpack_bf16x2(u.w); x[i][6] = f.x; x[i][7] = f.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) mx = fmaxf(mx, x[i][j]);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int v = lane + 32 * i;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float e = (mode == 0) ? bf16r(__expf(bf16r(x[i][j] - mx))) : __expf(x[i][j] - mx);
        x[i][j] = e;
        sum += e;
      }
    }
  }
  sum = warp_sum(sum);
  if (mode == 0) sum = bf16r(sum);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int v = lane + 32 * i;
    if (v < nvec) {
      uint4 u;
      u.x = pack_bf16x2(x[i][0] / sum, x[i][1] / sum);
      u.y = pack_bf16x2(x[i][2] / sum, x[i][3] / sum);
      u.z = pack_bf16x2(x[i][4] / sum, x[i][5] / sum);
      u.w = pack_bf16x2(x[i][6] / sum, x[i][7] / sum);
      *reinterpret_cast<uint4*>(s + v * 8) = u;
    }
  }
}

// delta[(bo * nbi + bi) * out_rows + out_off + r] = sum_d dO[bo, bi, r, d] * O[bo, bi, r, d]   (fp32; warp per row)
// — the row term of the softmax backward, dS = P o (dP - rowsum(P o dP)), with rowsum(P o dP) = dO . O: it lets the dP
// GEMM's epilogue (LAPB_EPI_SOFTMAX_BWD) emit dS directly instead of a separate pass over P and dP.
__global__ void __launch_bounds__(256)
rowdot_kernel(const bf16* __restrict__ dO, const bf16* __restrict__ O, float* __restrict__ delta, int rows, int D, long ldd,
              long ldo, int nbi, long d_bs_i, long d_bs_o, long o_bs_i, long o_bs_o, long out_rows, long out_off,
              long total_rows) {
  const long gr = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (gr >= total_rows) return;
  const int lane = threadIdx.x & 31;
  const int r = (int)(gr % rows);
  const long batch = gr / rows;
  const long bi = batch % nbi, bo = batch / nbi;
  const bf16* a = dO + bo * d_bs_o + bi * d_bs_i + (long)r * ldd;
  const bf16* b = O + bo * o_bs_o + bi * o_bs_i + (long)r * ldo;
  float acc = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    const uint4 ua = *reinterpret_cast<const uint4*>(a + c);
    const uint4 ub = *reinterpret_cast<const uint4*>(b + c);
    float2 x, y;
    x = unpack_bf16x2(ua.x); y = unpack_bf16x2(ub.x); acc += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(ua.y); y = unpack_bf16x2(ub.y); acc += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(ua.z); y = unpack_bf16x2(ub.z); acc += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(ua.w); y = unpack_bf16x2(ub.w); acc += x.x * y.x + x.y * y.y;
  }
  acc = warp_sum(acc);
  if (lane == 0) delta[batch * out_rows + out_off + r] = acc;
}

}  // namespace lapb

using namespace lapb;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int lapb200_mask_build(const uint8_t* pm, const uint8_t* par, const uint8_t* pma, const uint8_t* sm,
                       const uint8_t* sar, uint32_t* bits, int32_t* positions, int64_t B, int64_t P, int64_t A,
                       int64_t W32, int64_t row_begin, int64_t infer_rows, lapb_stream_t s) {
  LAPB_REQUIRE(W32 * 32 >= P + A, "mask_build: W32 too small");
  LAPB_REQUIRE(A == 0 || (sm && sar), "mask_build: suffix masks missing");
  size_t smem = 4 * (size_t)(P + A) * sizeof(int);
  LAPB_REQUIRE(smem <= 48 * 1024, "mask_build: sequence too long (%ld tokens)", (long)(P + A));
  // enough row chunks to fill the GPU at small batch (inference: B = 1), one chunk per sample at training batch sizes
  long nrows = P + A - row_begin;
  long chunks = (2L * num_sms() + B - 1) / B;
  if (chunks > (nrows + 7) / 8) chunks = (nrows + 7) / 8;
  if (chunks < 1) chunks = 1;
  mask_build_kernel<<<dim3((unsigned)B, (unsigned)chunks), 256, smem, STREAM(s)>>>(
      pm, par, pma, sm, sar, bits, positions, (int)P, (int)A, (int)W32, (int)row_begin, (int)infer_rows);
  LAPB_LAUNCH_OK("mask_build");
  return 0;
}

int lapb200_rowdot(const void* dO, const void* O, float* delta, int64_t rows, int64_t D, int64_t ldd, int64_t ldo,
                   int64_t nbi, int64_t nbo, int64_t d_bs_i, int64_t d_bs_o, int64_t o_bs_i, int64_t o_bs_o, int64_t out_rows,
                   int64_t out_off, lapb_stream_t s) {
  LAPB_REQUIRE(D % 8 == 0 && ldd % 8 == 0 && ldo % 8 == 0 && d_bs_i % 8 == 0 && d_bs_o % 8 == 0 && o_bs_i % 8 == 0 &&
               o_bs_o % 8 == 0, "rowdot: D, leading dimensions and batch strides must be multiples of 8 elements");
  const long total = rows * nbi * nbo;
  if (total == 0) return 0;
  rowdot_kernel<<<cdiv(total, 8), 256, 0, STREAM(s)>>>((const bf16*)dO, (const bf16*)O, delta, (int)rows, (int)D, ldd, ldo,
                                                       (int)nbi, d_bs_i, d_bs_o, o_bs_i, o_bs_o, out_rows, out_off, total);
  LAPB_LAUNCH_OK("rowdot");
  return 0;
}

int lapb200_mask_expand(const uint32_t* bits, uint8_t* dense, int64_t rows, int64_t S, int64_t W32, lapb_stream_t s) {
  mask_expand_kernel<<<148 * 4, 256, 0, STREAM(s)>>>(bits, dense, rows, (int)S, (int)W32);
  LAPB_LAUNCH_OK("mask_expand");
  return 0;
}

int lapb200_attn_softmax_fwd(const float* S, const uint32_t* bits, void* P, int64_t B, int64_t rows_per_batch,
                             int64_t G, int64_t S_len, int64_t ld, int64_t W32, lapb_stream_t s) {
  LAPB_REQUIRE(ld % 32 == 0 && ld <= 1024, "attn_softmax: ld must be a multiple of 32 and <= 1024 (got %ld)", (long)ld);
  LAPB_REQUIRE(rows_per_batch % G == 0, "attn_softmax: rows_per_batch %% G != 0");
  long total = B * rows_per_batch;
  attn_softmax_fwd_kernel<<<cdiv(total, 8), 256, 0, STREAM(s)>>>(S, bits, (bf16*)P, rows_per_batch, (int)G,
                                                                (int)S_len, (int)ld, (int)W32, total);
  LAPB_LAUNCH_OK("attn_softmax_fwd");
  return 0;
}

int lapb200_softmax_bwd(const void* P, const void* dP, void* dS, int64_t rows, int64_t ld, lapb_stream_t s) {
  LAPB_REQUIRE(ld % 2 == 0, "softmax_bwd: ld must be even");
  softmax_bwd_kernel<<<cdiv(rows, 8), 256, 0, STREAM(s)>>>((const bf16*)P, (const bf16*)dP, (bf16*)dS, (int)ld, rows);
  LAPB_LAUNCH_OK("softmax_bwd");
  return 0;
}

int lapb200_vit_softmax_fwd(void* S, int64_t rows, int64_t n, int64_t ld, int64_t mode, lapb_stream_t s) {
  LAPB_REQUIRE(n % 8 == 0 && ld % 8 == 0 && n <= 1024, "vit_softmax: n, ld must be multiples of 8 and n <= 1024");
  if (n <= 256)
    vit_softmax_fwd_kernel<1><<<cdiv(rows, 8), 256, 0, STREAM(s)>>>((bf16*)S, (int)n, (int)ld, rows, (int)mode);
  else if (n <= 512)
    vit_softmax_fwd_kernel<2><<<cdiv(rows, 8), 256, 0, STREAM(s)>>>((bf16*)S, (int)n, (int)ld, rows, (int)mode);
  else
    vit_softmax_fwd_kernel<4><<<cdiv(rows, 8), 256, 0, STREAM(s)>>>((bf16*)S, (int)n, (int)ld, rows, (int)mode);
  LAPB_LAUNCH_OK("vit_softmax_fwd");
  return 0;
}

}  // extern "C"
