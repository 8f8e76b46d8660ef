// K4-K8, K10 — HBM-bound kernels of the LAP hot path: casts, patchify, small fp32 GEMM, LayerNorm, (ada)RMSNorm,
// RoPE, GeGLU/GELU backward, column sums, embedding gather/scatter, gated-residual backward, suffix embedding.
//
// All of these stream their operands once with 128-bit loads/stores (8 bf16 or 4 fp32 per thread per access) and
// reduce with warp shuffles; the roofline that bounds them is HBM bandwidth.  C-ABI convention for this file:
// every integer argument is int64_t, every scalar float is `float`, last argument is the stream.
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"
#include <algorithm>

namespace lapb {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void ld8(const bf16* p, float (&x)[8]) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = unpack_bf16x2(v.x); x[0] = f.x; x[1] = f.y;
  f = unpack_bf16x2(v.y); x[2] = f.x; x[3] = f.y;
  f = unpack_bf16x2(v.z); x[4] = f.x; x[5] = f.y;
  f = unpack_bf16x2(v.w); x[6] = f.x; x[7] = f.y;
}
__device__ __forceinline__ void st8(bf16* p, const float (&x)[8]) {
  uint4 v;
  v.x = pack_bf16x2(x[0], x[1]); v.y = pack_bf16x2(x[2], x[3]);
  v.z = pack_bf16x2(x[4], x[5]); v.w = pack_bf16x2(x[6], x[7]);
  *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ void ld8f(const float* p, float (&x)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}

// ------------------------------------------------------------------------------------------------
// casts
// ------------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long n) {
  long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  long stride = (long)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float x[8];
    ld8f(src + i, x);
    st8(dst + i, x);
  }
  // tail (n % 8), handled by the first threads of the grid
  long tail0 = n & ~7L;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n - tail0) dst[tail0 + t] = __float2bfloat16_rn(src[tail0 + t]);
}

// dst[v, 0:D] = bf16(src[v]), dst[v, D:2D] = bf16(src[v] - float(hi)): the fp32 embedding table as a hi/lo bf16 pair so
// that the LM-head GEMM keeps ~16 mantissa bits of the fp32 table (gemma.py:153-154 multiplies by the fp32 table).
__global__ void split_hi_lo_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long rows, long D) {
  long nvec = rows * (D / 8);
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
    long r = v / (D / 8), c = (v % (D / 8)) * 8;
    float x[8], hi[8], lo[8];
    ld8f(src + r * D + c, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = bf16r(x[j]);
      lo[j] = x[j] - hi[j];
    }
    st8(dst + r * 2 * D + c, hi);
    st8(dst + r * 2 * D + D + c, lo);
  }
}

// ------------------------------------------------------------------------------------------------
// patchify: images (fp32 [-1,1] or uint8) -> fp32 patch rows [B*C*gh*gw, ps*ps*3], row order (b, cam, i, j),
// column order (p, q, ch) == the [ps,ps,3,width] conv kernel flattened (siglip.py:216-223).
// uint8 input applies Observation.from_dict's u8/255*2-1 (OP/models/model.py:116-118) on the fly.
// ------------------------------------------------------------------------------------------------
struct ImgPtrs {
  const void* p[4];
};
template <bool U8>
__global__ void patchify_kernel(ImgPtrs imgs, float* __restrict__ out, bf16* __restrict__ out_hi,
                                bf16* __restrict__ out_lo, int pk_pad, int B, int C, int H, int W, int ps) {
  int gh = H / ps, gw = W / ps, np = gh * gw, pk = ps * ps * 3;
  long total = (long)B * C * np * pk;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int col = idx % pk;
    long row = idx / pk;
    int t = row % np;
    long bc = row / np;
    int cam = bc % C, b = bc / C;
    int i = t / gw, j = t % gw;
    int ch = col % 3, pq = col / 3, q = pq % ps, p = pq / ps;
    long off = (((long)b * H + (i * ps + p)) * W + (j * ps + q)) * 3 + ch;
    float v;
    if (U8)
      v = (float)reinterpret_cast<const uint8_t*>(imgs.p[cam])[off] / 255.0f * 2.0f - 1.0f;
    else
      v = reinterpret_cast<const float*>(imgs.p[cam])[off];
    out[idx] = v;
    if (out_hi) {  // bf16 hi/lo split of the patch matrix (tensor-core weight gradient of the fp32 conv)
      float hi = bf16r(v);
      out_hi[row * pk_pad + col] = __float2bfloat16_rn(hi);
      out_lo[row * pk_pad + col] = __float2bfloat16_rn(v - hi);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// small fp32 GEMM on CUDA cores with generic strides (fp32 layers of the reference: patch conv, action_in_proj,
// time MLP, action_out_proj — SURVEY Appendix A.1/A.6/A.7 keeps them in fp32).
//   C[m,n] (+)= sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk]  (+ bias[n]) (+ table[(m % table_rows)*N + n])
// ------------------------------------------------------------------------------------------------
template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(const TA* __restrict__ A, const TB* __restrict__ B, void* __restrict__ Cv, int M, int N, int K, long sam,
             long sak, long sbn, long sbk, long ldc, const float* __restrict__ bias, const float* __restrict__ table,
             int table_rows, int c_bf16, int accumulate) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  // software pipeline: the global loads of K block k+1 are in flight (registers) while block k is multiplied out of
  // shared memory — these GEMMs are tiny (M <= 64 rows x K = 1024, or K = 588) and were bound by the exposed load latency
  // of every K block (~3 us each)
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * 256;
      int mm, kk;
      // pick the faster-varying index by stride so global loads coalesce where possible
      if (sak == 1) { kk = i % TK; mm = i / TK; } else { mm = i % TM; kk = i / TM; }
      int m = m0 + mm, k = k0 + kk;
      ra[q] = (m < M && k < K) ? (float)A[m * sam + k * sak] : 0.f;
      int nn;
      if (sbk == 1) { kk = i % TK; nn = i / TK; } else { nn = i % TN; kk = i / TN; }
      int n = n0 + nn;
      k = k0 + kk;
      rb[q] = (n < N && k < K) ? (float)B[n * sbn + k * sbk] : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * 256;
      int mm, kk;
      if (sak == 1) { kk = i % TK; mm = i / TK; } else { mm = i % TM; kk = i / TM; }
      As[kk][mm] = ra[q];
      int nn;
      if (sbk == 1) { kk = i % TK; nn = i / TK; } else { nn = i % TN; kk = i / TN; }
      Bs[kk][nn] = rb[q];
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += TK) {
    stash();
    __syncthreads();
    if (k0 + TK < K) fetch(k0 + TK);
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (table) v += table[(long)(m % table_rows) * N + n];
      if (c_bf16) {
        reinterpret_cast<bf16*>(Cv)[(long)m * ldc + n] = __float2bfloat16_rn(v);
      } else {
        float* c = reinterpret_cast<float*>(Cv) + (long)m * ldc + n;
        *c = accumulate ? (*c + v) : v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (flax nn.LayerNorm(dtype=bf16): fp32 stats, fast variance, eps 1e-6) — siglip.py:87,98,161
// one CTA (128 threads) per row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
layernorm_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ bias,
                     bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int W) {
  __shared__ float red[32];
  long row = blockIdx.x;
  const bf16* xr = x + row * W;
  float s = 0.f, s2 = 0.f;
  for (int c = threadIdx.x * 8; c < W; c += 128 * 8) {
    float v[8];
    ld8(xr + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += v[j]; s2 += v[j] * v[j]; }
  }
  s = block_sum(s, red);
  s2 = block_sum(s2, red);
  float mean = s / W;
  float var = fmaxf(s2 / W - mean * mean, 0.f);
  float rstd = rsqrtf(var + 1e-6f);
  if (threadIdx.x == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  for (int c = threadIdx.x * 8; c < W; c += 128 * 8) {
    float v[8], sc[8], bi[8], o[8];
    ld8(xr + c, v);
    ld8f(scale + c, sc);
    ld8f(bias + c, bi);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * (rstd * sc[j]) + bi[j];
    st8(y + row * W + c, o);
  }
}

// ------------------------------------------------------------------------------------------------
// Norm backward (LayerNorm: siglip.py:87,98,161; plain RMSNorm: gemma.py:112-131).
//   LN : dx = dres + rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*scale, xhat = (x-mean)*rstd;
//        dscale += sum_rows dy*xhat; dbias += sum_rows dy
//   RMS: dx[src] = dres[src] + rstd*(g - xhat*mean(g*xhat)), g = dy*(1+scale), xhat = x*rstd; dscale += sum dy*xhat
// One persistent CTA per SM, split into SLOTS independent row slots of TPB threads (one 16-byte vector per thread,
// width <= 2048).  Each slot walks its rows through a 4-deep shared-memory ring filled by 1-D TMA bulk copies that
// the slot leader issues four rows ahead, so every element crosses HBM once and the bytes in flight per SM (~190 KB)
// are set by the ring, not by registers.  A slot synchronises on its own named barrier once per row (row reduction +
// "stage consumed").  The per-column dscale/dbias partial sums stay in registers for the whole kernel; at the end
// the slots of a CTA are combined in shared memory and each CTA issues width/4 vector reductions
// (red.global.add.v4.f32): scalar atomics from ~700 CTAs were measured to cost more than the whole streaming pass.
// Everything the row loop touches is a compile-time constant (TPB, SLOTS, ring offsets) to keep it ~150 instructions.
// ------------------------------------------------------------------------------------------------
struct NormBwdArgs {
  const bf16* dy; long lddy;
  const bf16* x; long ldx;
  const long* row_idx;      // RMS only (optional): source/destination row of x, dres, dx
  const float* scale;
  const float* mean;        // LN only
  const float* rstd;
  const bf16* dres;         // optional
  bf16* dx;
  float* dscale;
  float* dbias;             // LN only
  long M;
  int D;
};
constexpr int NORM_STAGES = 4, NORM_R = 2;  // ring depth; rows per slot iteration (ILP across independent rows)
constexpr int norm_slots(int tpb) { return (512 / tpb) < 8 ? (512 / tpb) : 8; }
constexpr size_t norm_smem(int tpb) {
  return (size_t)norm_slots(tpb) * NORM_STAGES * NORM_R * 3 * tpb * 8 * sizeof(bf16);
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
  v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}

template <int TPB, bool LN>
__global__ void __launch_bounds__(norm_slots(TPB) * TPB, 1) norm_bwd_kernel(const NormBwdArgs a) {
  constexpr int SLOTS = norm_slots(TPB), ROW = TPB * 8, R = NORM_R, STAGE = R * 3 * ROW;
  extern __shared__ __align__(128) unsigned char ring_raw[];  // [SLOTS][NORM_STAGES][R][3][ROW] bf16
  __shared__ uint64_t full[SLOTS][NORM_STAGES];
  __shared__ __align__(16) float red[SLOTS][2][R][2][8];
  const int slot = threadIdx.x / TPB, t = threadIdx.x % TPB;
  bf16* ring = reinterpret_cast<bf16*>(ring_raw) + slot * (NORM_STAGES * STAGE);
  const int c = t * 8;
  const bool act = c < a.D;
  const bool has_res = a.dres != nullptr;
  const uint32_t row_bytes = (uint32_t)a.D * 2;
  const long stride = (long)gridDim.x * SLOTS * R;
  auto issue = [&](int stage, long row0) {  // rows row0 .. row0+R-1 (those < M)
    uint64_t* bar = &full[slot][stage];
    const int nrows = (int)(a.M - row0 < R ? a.M - row0 : R);
    mbar_expect_tx(bar, nrows * (has_res ? 3 : 2) * row_bytes);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r < nrows) {
        const long row = row0 + r;
        const long src = (!LN && a.row_idx) ? a.row_idx[row] : row;
        bf16* dst = ring + stage * STAGE + r * (3 * ROW);
        bulk_load_1d(dst, a.dy + row * a.lddy, row_bytes, bar);
        bulk_load_1d(dst + ROW, a.x + src * a.ldx, row_bytes, bar);
        if (has_res) bulk_load_1d(dst + 2 * ROW, a.dres + src * a.ldx, row_bytes, bar);
      }
    }
  };
  long row0 = ((long)blockIdx.x * SLOTS + slot) * R;
  if (t == 0) {
#pragma unroll
    for (int s = 0; s < NORM_STAGES; ++s) mbar_init(&full[slot][s], 1);
    fence_barrier_init();
#pragma unroll
    for (int s = 0; s < NORM_STAGES; ++s) {
      const long r = row0 + s * stride;
      if (r < a.M) issue(s, r);
    }
  }
  for (int i = t; i < 2 * R * 2 * 8; i += TPB) (&red[slot][0][0][0][0])[i] = 0.f;
  float ds[8], db[8], sc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ds[j] = 0.f; db[j] = 0.f; sc[j] = 0.f; }
  if (act) {
    ld8f(a.scale + c, sc);
    if (!LN) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sc[j] += 1.0f;
    }
  }
  __syncthreads();
  const float invD = 1.0f / a.D;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  int s = 0, ph = 0, par = 0;
  float nmu[R], nrs[R];
  long nsrc[R];
  auto fetch_stats = [&](long base) {  // per-row scalars of the NEXT iteration (consumed one iteration later)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long row = base + r < a.M ? base + r : a.M - 1;
      nrs[r] = a.rstd[row];
      nmu[r] = LN ? a.mean[row] : 0.f;
      nsrc[r] = (!LN && a.row_idx) ? a.row_idx[row] : row;
    }
  };
  if (row0 < a.M) fetch_stats(row0);
  for (; row0 < a.M; row0 += stride) {
    float mu[R], rs[R];
    long src[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { mu[r] = nmu[r]; rs[r] = nrs[r]; src[r] = nsrc[r]; }
    if (row0 + stride < a.M) fetch_stats(row0 + stride);
    mbar_wait(&full[slot][s], ph);
    float g[R][8], xh[R][8], sg[R], sgx[R];
    uint4 vr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool on = act && (row0 + r < a.M);  // columns past the width / rows past M hold stale ring bytes
      const bf16* st = ring + s * STAGE + r * (3 * ROW) + c;
      const uint4 vd = *reinterpret_cast<const uint4*>(st);
      const uint4 vx = *reinterpret_cast<const uint4*>(st + ROW);
      vr[r] = has_res ? *reinterpret_cast<const uint4*>(st + 2 * ROW) : zero4;
      unpack8(vd, g[r]);
      unpack8(vx, xh[r]);
      sg[r] = 0.f;
      sgx[r] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!on) { g[r][j] = 0.f; xh[r][j] = 0.f; }
        xh[r][j] = LN ? (xh[r][j] - mu[r]) * rs[r] : xh[r][j] * rs[r];
        if (!on) xh[r][j] = 0.f;
        ds[j] = fmaf(g[r][j], xh[r][j], ds[j]);
        if (LN) db[j] += g[r][j];
        g[r][j] *= sc[j];
        if (LN) sg[r] += g[r][j];
        sgx[r] = fmaf(g[r][j], xh[r][j], sgx[r]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        sgx[r] += __shfl_xor_sync(0xffffffffu, sgx[r], o);
        if (LN) sg[r] += __shfl_xor_sync(0xffffffffu, sg[r], o);
      }
    }
    if ((t & 31) == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        red[slot][par][r][0][t >> 5] = sgx[r];
        if (LN) red[slot][par][r][1][t >> 5] = sg[r];
      }
    }
    named_bar_sync(1 + slot, TPB);  // also: every thread of the slot has read stage s -> it may be refilled
    if (t == 0) {
      const long r = row0 + NORM_STAGES * stride;
      if (r < a.M) issue(s, r);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 r0 = *reinterpret_cast<const float4*>(&red[slot][par][r][0][0]);
      const float4 r1 = *reinterpret_cast<const float4*>(&red[slot][par][r][0][4]);
      sgx[r] = (((r0.x + r0.y) + (r0.z + r0.w)) + ((r1.x + r1.y) + (r1.z + r1.w))) * invD;
      if (LN) {
        const float4 q0 = *reinterpret_cast<const float4*>(&red[slot][par][r][1][0]);
        const float4 q1 = *reinterpret_cast<const float4*>(&red[slot][par][r][1][4]);
        sg[r] = (((q0.x + q0.y) + (q0.z + q0.w)) + ((q1.x + q1.y) + (q1.z + q1.w))) * invD;
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (act && row0 + r < a.M) {
        float rr[8];
        unpack8(vr[r], rr);
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float o0 = fmaf(-xh[r][j], sgx[r], g[r][j]), o1 = fmaf(-xh[r][j + 1], sgx[r], g[r][j + 1]);
          if (LN) { o0 -= sg[r]; o1 -= sg[r]; }
          ow[j >> 1] = pack_bf16x2(fmaf(rs[r], o0, rr[j]), fmaf(rs[r], o1, rr[j + 1]));
        }
        *reinterpret_cast<uint4*>(a.dx + src[r] * a.ldx + c) = o;
      }
    }
    par ^= 1;
    if (++s == NORM_STAGES) { s = 0; ph ^= 1; }
  }
  // combine the slots of this CTA (the ring is idle now: every issued copy has been waited on)
  __syncthreads();
  float* acc = reinterpret_cast<float*>(ring_raw);  // [SLOTS][2][ROW]
  if (act) {
    *reinterpret_cast<float4*>(acc + (slot * 2 + 0) * ROW + c) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    *reinterpret_cast<float4*>(acc + (slot * 2 + 0) * ROW + c + 4) = make_float4(ds[4], ds[5], ds[6], ds[7]);
    if (LN) {
      *reinterpret_cast<float4*>(acc + (slot * 2 + 1) * ROW + c) = make_float4(db[0], db[1], db[2], db[3]);
      *reinterpret_cast<float4*>(acc + (slot * 2 + 1) * ROW + c + 4) = make_float4(db[4], db[5], db[6], db[7]);
    }
  }
  __syncthreads();
  const int nq = a.D / 4;
  for (int i = threadIdx.x; i < (LN ? 2 : 1) * nq; i += SLOTS * TPB) {
    const int which = i >= nq ? 1 : 0, c4 = (i - which * nq) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
      const float4 u = *reinterpret_cast<const float4*>(acc + (q * 2 + which) * ROW + c4);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    red_add_v4((which ? a.dbias : a.dscale) + c4, v.x, v.y, v.z, v.w);
  }
}

template <bool LN>
static int launch_norm_bwd(const NormBwdArgs& a, cudaStream_t stream) {
  const int tpb = (a.D / 8 + 31) / 32 * 32;
#define NORM_BWD_CASE(T)                                                                                          \
  case T: {                                                                                                       \
    static bool attr_set = false;                                                                                 \
    if (!attr_set) {                                                                                              \
      LAPB_CUDA_OK(cudaFuncSetAttribute(norm_bwd_kernel<T, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                        (int)norm_smem(T)));                                                      \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    constexpr int rows_per_cta = norm_slots(T) * NORM_R;                                                         \
    const int grid = (int)std::min<long>((a.M + rows_per_cta - 1) / rows_per_cta, (long)num_sms());              \
    norm_bwd_kernel<T, LN><<<grid, norm_slots(T) * T, norm_smem(T), stream>>>(a);                                 \
    break;                                                                                                        \
  }
  switch (tpb) {
    NORM_BWD_CASE(32)
    NORM_BWD_CASE(64)
    NORM_BWD_CASE(96)
    NORM_BWD_CASE(128)
    NORM_BWD_CASE(160)
    NORM_BWD_CASE(192)
    NORM_BWD_CASE(224)
    NORM_BWD_CASE(256)
    default:
      return set_error(-1, "norm_bwd: width %d not supported (must be <= 2048)", a.D);
  }
#undef NORM_BWD_CASE
  return 0;
}

// ------------------------------------------------------------------------------------------------
// RMSNorm / adaptive RMSNorm — gemma.py:112-131
//   plain:    y = bf16( x*rsqrt(mean(x^2)+1e-6) * (1+scale) )
//   adaptive: y = bf16( x*rstd * bf16(1+scale_b) + shift_b ), (scale_b, shift_b, gate_b) = chunks of mod[b]
// one CTA (128 threads) per row; D <= 2048... generic loop.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
rmsnorm_fwd_kernel(const bf16* __restrict__ x, long ldx, const long* __restrict__ row_idx,
                   const float* __restrict__ scale, const bf16* __restrict__ mod, long ldmod, int rows_per_sample,
                   bf16* __restrict__ y, long ldy, int dup, float* __restrict__ rstd_out, int D) {
  __shared__ float red[32];
  long row = blockIdx.x;
  long src = row_idx ? row_idx[row] : row;
  const bf16* xr = x + src * ldx;
  float s2 = 0.f;
  for (int c = threadIdx.x * 8; c < D; c += 128 * 8) {
    float v[8];
    ld8(xr + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s2 += v[j] * v[j];
  }
  s2 = block_sum(s2, red);
  float rstd = rsqrtf(s2 / D + 1e-6f);
  if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
  const bf16* m = mod ? mod + (row / rows_per_sample) * ldmod : nullptr;
  for (int c = threadIdx.x * 8; c < D; c += 128 * 8) {
    float v[8], o[8];
    ld8(xr + c, v);
    if (m) {
      float sc[8], sh[8];
      ld8(m + c, sc);
      ld8(m + D + c, sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[j] * rstd) * bf16r(1.0f + sc[j]) + sh[j];
    } else {
      float sc[8];
      ld8f(scale + c, sc);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[j] * rstd) * (1.0f + sc[j]);
    }
    st8(y + row * ldy + c, o);
    if (dup) st8(y + row * ldy + D + c, o);  // [y | y] for the split-table LM head
  }
}

// Finalisation of a deterministic split-K GEMM fused with the following normalisation (batch-1 inference: the
// M = 512 / 692-row down and fc2 projections have too few output tiles for 148 SMs, so they run as k_splits slabs):
//   t  = bf16( sum_s acc[s][row, :] )                      (slabs summed in fixed order)
//   t  = bf16( t + bf16(bias) )                            (flax Dense bias, SigLIP only)
//   x2 = bf16( resid + t )             -> xout             (the residual add of LAPB_EPI_RESID, same rounding points)
//   y  = LayerNorm(x2) or RMSNorm(x2)  -> y (optional)     (the next block's pre-norm; statistics in fp32)
// One CTA of 128 threads per row, the row stays in registers (D <= 3072).
template <bool LN>
__global__ void __launch_bounds__(128)
resid_norm_fwd_kernel(const bf16* __restrict__ resid, const float* __restrict__ acc, int nsplit, long slab_stride,
                      const float* __restrict__ bias, bf16* __restrict__ xout, const float* __restrict__ scale,
                      const float* __restrict__ nbias, bf16* __restrict__ y, float* __restrict__ mean_out,
                      float* __restrict__ rstd_out, int D) {
  __shared__ float red[32];
  const long row = blockIdx.x;
  float v[3][8];
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int c = threadIdx.x * 8 + k * 1024;
    if (c < D) {
      float t[8], r[8];
      const float* ap = acc + row * D + c;
      ld8f(ap, t);
      for (int sp = 1; sp < nsplit; ++sp) {
        float u[8];
        ld8f(ap + sp * slab_stride, u);
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] += u[j];
      }
      ld8(resid + row * D + c, r);
      if (bias) {
        float b[8];
        ld8f(bias + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = bf16r(t[j]) + bf16r(b[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[k][j] = bf16r(r[j] + bf16r(t[j]));
        s += v[k][j];
        s2 += v[k][j] * v[k][j];
      }
      st8(xout + row * D + c, v[k]);
    }
  }
  if (y == nullptr) return;
  s2 = block_sum(s2, red);
  float mean = 0.f, rstd;
  if (LN) {
    s = block_sum(s, red);
    mean = s / D;
    rstd = rsqrtf(fmaxf(s2 / D - mean * mean, 0.f) + 1e-6f);
    if (threadIdx.x == 0 && mean_out) mean_out[row] = mean;
  } else {
    rstd = rsqrtf(s2 / D + 1e-6f);
  }
  if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int c = threadIdx.x * 8 + k * 1024;
    if (c < D) {
      float sc[8], o[8];
      ld8f(scale + c, sc);
      if (LN) {
        float bi[8];
        ld8f(nbias + c, bi);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean) * (rstd * sc[j]) + bi[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[k][j] * rstd) * (1.0f + sc[j]);
      }
      st8(y + row * D + c, o);
    }
  }
}

// Warp-per-row forward norms for contiguous rows of width <= 2048 (the training shapes: 22144 x 2048 RMSNorm, 16384 x 1152
// LayerNorm).  A lane keeps its share of the row (<= 8 vectors of 16 bytes) in registers, so every element crosses HBM
// once each way; the only synchronisation is the warp shuffle of the row statistics.  The CTA-per-row kernels above
// launch 22144 CTAs of 128 threads for 4 KB rows: 52 % / 20 % of the HBM peak; these reach the streaming rate.
template <bool LN>
__global__ void __launch_bounds__(256)
norm_fwd_warp_kernel(const bf16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ bias,
                     bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long M, int D) {
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (long)gridDim.x * 8;
  const int nvec = D >> 3;
  for (long row = warp; row < M; row += nwarps) {
    const bf16* xr = x + row * D;
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = lane + 32 * k;
      v[k] = (c < nvec) ? *reinterpret_cast<const uint4*>(xr + 8 * c) : make_uint4(0, 0, 0, 0);
    }
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float f[8];
      unpack8(v[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += f[j];
        s2 += f[j] * f[j];
      }
    }
    s2 = warp_sum(s2);
    float mean = 0.f, rstd;
    if (LN) {
      s = warp_sum(s);
      mean = s / D;
      const float var = fmaxf(s2 / D - mean * mean, 0.f);
      rstd = rsqrtf(var + 1e-6f);
      if (lane == 0) mean_out[row] = mean;
    } else {
      rstd = rsqrtf(s2 / D + 1e-6f);
    }
    if (lane == 0 && rstd_out) rstd_out[row] = rstd;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = lane + 32 * k;
      if (c < nvec) {
        float f[8], sc[8], o[8];
        unpack8(v[k], f);
        ld8f(scale + 8 * c, sc);
        if (LN) {
          float bi[8];
          ld8f(bias + 8 * c, bi);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (f[j] - mean) * (rstd * sc[j]) + bi[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (f[j] * rstd) * (1.0f + sc[j]);
        }
        st8(y + row * D + 8 * c, o);
      }
    }
  }
}

// adaptive RMSNorm backward, one CTA per sample (rows_per_sample rows):
//   dx = dres + rstd*(g - xhat*mean(g*xhat)), g = dy*bf16(1+scale_b)
//   dmod[b, 0:D] (+)= sum_rows dy*xhat ; dmod[b, D:2D] (+)= sum_rows dy      (bf16 out; gate part written elsewhere)
__global__ void __launch_bounds__(128)
ada_rmsnorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ mod,
                       long ldmod, const float* __restrict__ rstd, const bf16* __restrict__ dres,
                       bf16* __restrict__ dx, bf16* __restrict__ dmod, long lddmod, int rows_per_sample, int D) {
  __shared__ float red[32];
  constexpr int MAXV = 2;
  float dsc[MAXV][8] = {}, dsh[MAXV][8] = {};
  long b = blockIdx.x;
  const bf16* m = mod + b * ldmod;
  for (int rr_ = 0; rr_ < rows_per_sample; ++rr_) {
    long row = b * rows_per_sample + rr_;
    float rs = rstd[row];
    float g[MAXV][8], xh[MAXV][8];
    float sgx = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = (threadIdx.x + i * 128) * 8;
      if (c < D) {
        float d[8], xv[8], sc[8];
        ld8(dy + row * D + c, d);
        ld8(x + row * D + c, xv);
        ld8(m + c, sc);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = xv[j] * rs;
          g[i][j] = d[j] * bf16r(1.0f + sc[j]);
          sgx += g[i][j] * xh[i][j];
          dsc[i][j] += d[j] * xh[i][j];
          dsh[i][j] += d[j];
        }
      }
    }
    sgx = block_sum(sgx, red) / D;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = (threadIdx.x + i * 128) * 8;
      if (c < D) {
        float o[8], rr[8];
        if (dres) ld8(dres + row * D + c, rr);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = rs * (g[i][j] - xh[i][j] * sgx);
          if (dres) o[j] += rr[j];
        }
        st8(dx + row * D + c, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = (threadIdx.x + i * 128) * 8;
    if (c < D) {
      st8(dmod + b * lddmod + c, dsc[i]);
      st8(dmod + b * lddmod + D + c, dsh[i]);
    }
  }
}

// gated residual backward (gemma.py:583: x_out = x + y*gate), one CTA per sample:
//   dy = dxo * gate (bf16);  dgate[b] = sum_rows dxo * y  -> dmod[b, 2D:3D]
__global__ void __launch_bounds__(128)
gated_bwd_kernel(const bf16* __restrict__ dxo, const bf16* __restrict__ y, const bf16* __restrict__ gate, long ldg,
                 bf16* __restrict__ dy, bf16* __restrict__ dgate, long lddg, int rows_per_sample, int D) {
  long b = blockIdx.x;
  for (int c = threadIdx.x * 8; c < D; c += 128 * 8) {
    float gt[8], acc[8] = {};
    ld8(gate + b * ldg + c, gt);
    for (int r = 0; r < rows_per_sample; ++r) {
      long row = b * rows_per_sample + r;
      float d[8], yv[8], o[8];
      ld8(dxo + row * D + c, d);
      ld8(y + row * D + c, yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = d[j] * gt[j];
        acc[j] += d[j] * yv[j];
      }
      st8(dy + row * D + c, o);
    }
    st8(dgate + b * lddg + c, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// RoPE (gemma.py:215-218,548-564): q <- bf16(rope(q)) * hd^-0.5, k <- bf16(rope(k)); gathers the fused QKV
// projection of both experts into attention layout:
//   Q [B, T, NH, HD]   K,V [B, Tpad, HD]   (T = P + A; rows [0,P) from the prefix expert, [P,T) from the suffix expert)
// one CTA per (b, t); thread handles 8 consecutive dims of the half-split pairs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
rope_fwd_kernel(const bf16* __restrict__ qkv0, const bf16* __restrict__ qkv1, const int* __restrict__ positions,
                const float* __restrict__ timescale, bf16* __restrict__ Q, bf16* __restrict__ Kc,
                bf16* __restrict__ Vc, int P, int A, int Tpad, int NH, int HD, int t_begin, int Tpos, float qscale) {
  int T = P + A;
  int nt = T - t_begin;
  int b = blockIdx.x / nt, t = t_begin + blockIdx.x % nt;
  int ld = (NH + 2) * HD;
  const bf16* src = (t < P) ? qkv0 + ((long)b * P + t) * ld : qkv1 + ((long)b * A + (t - P)) * ld;
  float pos = (float)positions[(long)b * Tpos + (t - t_begin)];
  int half = HD / 2;
  int nvec_head = half / 8;               // vectors per half head
  int total = (NH + 1) * nvec_head;       // q heads + 1 k head
  for (int v = threadIdx.x; v < total; v += 128) {
    int h = v / nvec_head, c = (v % nvec_head) * 8;
    const bf16* s = src + (long)h * HD;  // h == NH -> K (immediately after the NH query heads)
    float x1[8], x2[8], o1[8], o2[8];
    ld8(s + c, x1);
    ld8(s + half + c, x2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float rad = pos / timescale[c + j];
      float sn, cs;
      sincosf(rad, &sn, &cs);
      o1[j] = bf16r(x1[j] * cs - x2[j] * sn);
      o2[j] = bf16r(x2[j] * cs + x1[j] * sn);
      if (h < NH) { o1[j] *= qscale; o2[j] *= qscale; }
    }
    bf16* d = (h < NH) ? Q + (((long)b * nt + (t - t_begin)) * NH + h) * HD : Kc + ((long)b * Tpad + t) * HD;
    st8(d + c, o1);
    st8(d + half + c, o2);
  }
  // V: plain copy
  for (int c = threadIdx.x * 8; c < HD; c += 128 * 8) {
    uint4 v = *reinterpret_cast<const uint4*>(src + (long)(NH + 1) * HD + c);
    *reinterpret_cast<uint4*>(Vc + ((long)b * Tpad + t) * HD + c) = v;
  }
}

// inverse: d(qkv) from dQ [B,T,NH,HD], dK,dV [B,Tpad,HD]
__global__ void __launch_bounds__(128)
rope_bwd_kernel(const bf16* __restrict__ dQ, const bf16* __restrict__ dK, const bf16* __restrict__ dV,
                const int* __restrict__ positions, const float* __restrict__ timescale, bf16* __restrict__ dqkv0,
                bf16* __restrict__ dqkv1, int P, int A, int Tpad, int NH, int HD, float qscale) {
  int T = P + A;
  int b = blockIdx.x / T, t = blockIdx.x % T;
  int ld = (NH + 2) * HD;
  bf16* dst = (t < P) ? dqkv0 + ((long)b * P + t) * ld : dqkv1 + ((long)b * A + (t - P)) * ld;
  float pos = (float)positions[(long)b * T + t];
  int half = HD / 2, nvec_head = half / 8, total = (NH + 1) * nvec_head;
  for (int v = threadIdx.x; v < total; v += 128) {
    int h = v / nvec_head, c = (v % nvec_head) * 8;
    const bf16* s = (h < NH) ? dQ + (((long)b * T + t) * NH + h) * HD : dK + ((long)b * Tpad + t) * HD;
    float g1[8], g2[8], o1[8], o2[8];
    ld8(s + c, g1);
    ld8(s + half + c, g2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float rad = pos / timescale[c + j];
      float sn, cs;
      sincosf(rad, &sn, &cs);
      float a = g1[j], bb = g2[j];
      if (h < NH) { a *= qscale; bb *= qscale; }
      // forward: o1 = x1 c - x2 s ; o2 = x2 c + x1 s  =>  dx1 = a c + b s ; dx2 = -a s + b c
      o1[j] = a * cs + bb * sn;
      o2[j] = -a * sn + bb * cs;
    }
    st8(dst + (long)h * HD + c, o1);
    st8(dst + (long)h * HD + half + c, o2);
  }
  for (int c = threadIdx.x * 8; c < HD; c += 128 * 8) {
    uint4 v = *reinterpret_cast<const uint4*>(dV + ((long)b * Tpad + t) * HD + c);
    *reinterpret_cast<uint4*>(dst + (long)(NH + 1) * HD + c) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// GeGLU backward (lora.py:124-142), in place:  in  dact=[dAct], gu=[g|u]   out  dact=[act], gu=[dg|du]
// WRITE_ACT = false: the caller kept `act` from the forward pass (one sixth of this kernel's HBM traffic saved)
// ------------------------------------------------------------------------------------------------
template <bool WRITE_ACT>
__global__ void geglu_bwd_kernel(bf16* __restrict__ dact, bf16* __restrict__ gu, long M, long F) {
  long nvec = M * (F / 8);
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
    long r = v / (F / 8), c = (v % (F / 8)) * 8;
    float d[8], g[8], u[8], act[8], dg[8], du[8];
    ld8(dact + r * F + c, d);
    ld8(gu + r * 2 * F + c, g);
    ld8(gu + r * 2 * F + F + c, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float ge = bf16r(gelu_tanh(g[j]));
      act[j] = ge * u[j];
      du[j] = d[j] * ge;
      dg[j] = d[j] * u[j] * gelu_tanh_grad(g[j]);
    }
    if (WRITE_ACT) st8(dact + r * F + c, act);
    st8(gu + r * 2 * F + c, dg);
    st8(gu + r * 2 * F + F + c, du);
  }
}

// GELU backward (siglip.py:71), in place on dh:  dh <- dh * gelu'(pre)
__global__ void gelu_bwd_kernel(bf16* __restrict__ dh, const bf16* __restrict__ pre, long n) {
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += (long)gridDim.x * blockDim.x * 8) {
    float d[8], p[8];
    ld8(dh + i, d);
    ld8(pre + i, p);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] *= gelu_tanh_grad(p[j]);
    st8(dh + i, d);
  }
}

// swish (nnx.swish) forward / backward on small fp32 tensors (time MLP, pi0.py:165-167)
__global__ void swish_fwd_kernel(const float* __restrict__ z, float* __restrict__ y, bf16* __restrict__ y_bf16,
                                 long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = z[i];
    float o = v / (1.0f + expf(-v));
    y[i] = o;
    if (y_bf16) y_bf16[i] = __float2bfloat16_rn(o);
  }
}
__global__ void swish_bwd_kernel(const float* __restrict__ z, const float* __restrict__ dy,
                                 const bf16* __restrict__ dy_bf16, float* __restrict__ dz, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = z[i];
    float s = 1.0f / (1.0f + expf(-v));
    float d = dy ? dy[i] : __bfloat162float(dy_bf16[i]);
    dz[i] = d * (s + v * s * (1.0f - s));
  }
}

// ------------------------------------------------------------------------------------------------
// column sums of a bf16 matrix: out[n] += sum_m X[m, n]  (bias / pos-embedding gradients)
// grid (ceil(N/256), row_chunks); each thread owns 2 adjacent columns... simple version: one column per thread
// ------------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const bf16* __restrict__ X, long ldx, float* __restrict__ out, long M, long N,
                              long rows_per_cta) {
  long n = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (n >= N) return;
  long r0 = (long)blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float a0 = 0.f, a1 = 0.f;
  for (long r = r0; r < r1; ++r) {
    float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(X + r * ldx + n));
    a0 += f.x;
    a1 += f.y;
  }
  atomicAdd(out + n, a0);
  if (n + 1 < N) atomicAdd(out + n + 1, a1);
}

// ------------------------------------------------------------------------------------------------
// embedding gather (gemma.py:148-151,446-448): X[b, off + j, :] = bf16(E[ids[b,j]] * sqrt(D)); scatter-add backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
embed_fwd_kernel(const int* __restrict__ ids, const float* __restrict__ E, bf16* __restrict__ X, int L, long row_off,
                 long rows_per_sample, int D, float scale) {
  long i = blockIdx.x;  // token index b*L + j
  long b = i / L, j = i % L;
  const float* e = E + (long)ids[i] * D;
  bf16* x = X + (b * rows_per_sample + row_off + j) * D;
  for (int c = threadIdx.x * 8; c < D; c += 128 * 8) {
    float v[8];
    ld8f(e + c, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= scale;
    st8(x + c, v);
  }
}
__global__ void __launch_bounds__(128)
embed_bwd_kernel(const int* __restrict__ ids, const bf16* __restrict__ dX, float* __restrict__ dE, int L, long row_off,
                 long rows_per_sample, int D, float scale) {
  long i = blockIdx.x;
  long b = i / L, j = i % L;
  float* e = dE + (long)ids[i] * D;
  const bf16* x = dX + (b * rows_per_sample + row_off + j) * D;
  for (int c = threadIdx.x * 8; c < D; c += 128 * 8) {
    float v[8];
    ld8(x + c, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(e + c + k, v[k] * scale);
  }
}

// dX[rows[i], :] = d[i, :]   (dX pre-zeroed; rows unique)
__global__ void __launch_bounds__(128)
scatter_rows_kernel(const bf16* __restrict__ d, const long* __restrict__ rows, bf16* __restrict__ dX, int D) {
  long i = blockIdx.x;
  for (int c = threadIdx.x * 8; c < D; c += 128 * 8)
    *reinterpret_cast<uint4*>(dX + rows[i] * D + c) = *reinterpret_cast<const uint4*>(d + i * D + c);
}

// ------------------------------------------------------------------------------------------------
// flow-matching suffix inputs (lap.py:193-197, pi0.py:47-63): x_t = t*noise + (1-t)*a ; u_t = noise - a ;
// time_emb[b] = [sin, cos](t_b * 2*pi / period_i), period_i = 4e-3 * 1000^(i/(W/2-1))
// ------------------------------------------------------------------------------------------------
__global__ void suffix_inputs_kernel(const float* __restrict__ actions, const float* __restrict__ noise,
                                     const float* __restrict__ time, float* __restrict__ x_t, float* __restrict__ u_t,
                                     float* __restrict__ time_emb, int B, int AD /*A*ad*/, int W) {
  int b = blockIdx.x;
  float t = time[b];
  if (actions) {
    for (int i = threadIdx.x; i < AD; i += blockDim.x) {
      float a = actions[(long)b * AD + i], n = noise[(long)b * AD + i];
      x_t[(long)b * AD + i] = t * n + (1.0f - t) * a;
      u_t[(long)b * AD + i] = n - a;
    }
  }
  int half = W / 2;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    float fraction = (half > 1) ? (float)i / (float)(half - 1) : 0.f;
    float period = 4e-3f * powf(4.0f / 4e-3f, fraction);
    float inp = t * (1.0f / period * 2.0f * 3.14159265358979323846f);
    float sn, cs;
    sincosf(inp, &sn, &cs);
    time_emb[(long)b * W + i] = sn;
    time_emb[(long)b * W + half + i] = cs;
  }
}

// x += dt * v  (Euler step, lap.py:667)
__global__ void axpy_kernel(float* __restrict__ x, const float* __restrict__ v, float dt, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    x[i] += dt * v[i];
}

static inline int grid_for(long work_items, int block, int max_blocks = 148 * 16) {
  long g = (work_items + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace lapb

using namespace lapb;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int lapb200_cast_f32_bf16(const float* src, void* dst, int64_t n, lapb_stream_t s) {
  cast_f32_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, STREAM(s)>>>(src, (bf16*)dst, n);
  LAPB_LAUNCH_OK("cast_f32_bf16");
  return 0;
}

int lapb200_split_hi_lo(const float* src, void* dst, int64_t rows, int64_t D, lapb_stream_t s) {
  LAPB_REQUIRE(D % 8 == 0, "split_hi_lo: D %% 8 != 0");
  split_hi_lo_kernel<<<grid_for(rows * D / 8, 256), 256, 0, STREAM(s)>>>(src, (bf16*)dst, rows, D);
  LAPB_LAUNCH_OK("split_hi_lo");
  return 0;
}

int lapb200_patchify(const void* img0, const void* img1, const void* img2, int64_t is_u8, float* out, void* out_hi,
                     void* out_lo, int64_t pk_pad, int64_t B, int64_t C, int64_t H, int64_t W, int64_t ps,
                     lapb_stream_t s) {
  LAPB_REQUIRE(C >= 1 && C <= 3, "patchify: 1..3 cameras");
  ImgPtrs p;
  p.p[0] = img0; p.p[1] = img1; p.p[2] = img2; p.p[3] = nullptr;
  long total = B * C * (H / ps) * (W / ps) * ps * ps * 3;
  if (is_u8)
    patchify_kernel<true><<<grid_for(total, 256), 256, 0, STREAM(s)>>>(p, out, (bf16*)out_hi, (bf16*)out_lo, (int)pk_pad, B, C, H, W, ps);
  else
    patchify_kernel<false><<<grid_for(total, 256), 256, 0, STREAM(s)>>>(p, out, (bf16*)out_hi, (bf16*)out_lo, (int)pk_pad, B, C, H, W, ps);
  LAPB_LAUNCH_OK("patchify");
  return 0;
}

// dtype flags: 0 = fp32, 1 = bf16
int lapb200_sgemm(const void* A, int64_t a_bf16, const void* B, int64_t b_bf16, void* C, int64_t c_bf16, int64_t M,
                  int64_t N, int64_t K, int64_t sam, int64_t sak, int64_t sbn, int64_t sbk, int64_t ldc,
                  const float* bias, const float* table, int64_t table_rows, int64_t accumulate, lapb_stream_t s) {
  dim3 grid(cdiv(N, 64), cdiv(M, 64));
  int tr = table_rows > 0 ? (int)table_rows : 1;
#define SGEMM_LAUNCH(TA, TB)                                                                                         \
  sgemm_kernel<TA, TB><<<grid, 256, 0, STREAM(s)>>>((const TA*)A, (const TB*)B, C, (int)M, (int)N, (int)K, sam, sak, \
                                                    sbn, sbk, ldc, bias, table, tr, (int)c_bf16, (int)accumulate)
  if (!a_bf16 && !b_bf16) SGEMM_LAUNCH(float, float);
  else if (a_bf16 && !b_bf16) SGEMM_LAUNCH(bf16, float);
  else if (!a_bf16 && b_bf16) SGEMM_LAUNCH(float, bf16);
  else SGEMM_LAUNCH(bf16, bf16);
#undef SGEMM_LAUNCH
  LAPB_LAUNCH_OK("sgemm");
  return 0;
}

int lapb200_layernorm_fwd(const void* x, const float* scale, const float* bias, void* y, float* mean, float* rstd,
                          int64_t M, int64_t W, lapb_stream_t s) {
  LAPB_REQUIRE(W % 8 == 0, "layernorm: W %% 8 != 0");
  if (W <= 2048 && M >= 1024) {
    norm_fwd_warp_kernel<true><<<grid_for(M, 8, 148 * 8), 256, 0, STREAM(s)>>>((const bf16*)x, scale, bias, (bf16*)y,
                                                                             mean, rstd, M, (int)W);
    LAPB_LAUNCH_OK("layernorm_fwd");
    return 0;
  }
  layernorm_fwd_kernel<<<(unsigned)M, 128, 0, STREAM(s)>>>((const bf16*)x, scale, bias, (bf16*)y, mean, rstd, (int)W);
  LAPB_LAUNCH_OK("layernorm_fwd");
  return 0;
}

int lapb200_layernorm_bwd(const void* dy, const void* x, const float* scale, const float* mean, const float* rstd,
                          const void* dres, void* dx, float* dscale, float* dbias, int64_t M, int64_t W,
                          lapb_stream_t s) {
  LAPB_REQUIRE(W % 8 == 0 && W <= 2048, "layernorm_bwd: W must be a multiple of 8 and <= 2048");
  NormBwdArgs a{(const bf16*)dy, W, (const bf16*)x, W, nullptr, scale, mean, rstd, (const bf16*)dres, (bf16*)dx, dscale,
                dbias, M, (int)W};
  if (M <= 0) return 0;
  if (int rc = launch_norm_bwd<true>(a, STREAM(s))) return rc;
  LAPB_LAUNCH_OK("layernorm_bwd");
  return 0;
}

int lapb200_resid_norm_fwd(const void* resid, const float* acc, int64_t nsplit, int64_t slab_stride, const float* bias,
                           void* xout, int64_t layernorm, const float* scale, const float* nbias, void* y, float* mean,
                           float* rstd, int64_t M, int64_t D, lapb_stream_t s) {
  LAPB_REQUIRE(D % 8 == 0 && D <= 3072 && nsplit >= 1, "resid_norm_fwd: D must be a multiple of 8 and <= 3072");
  LAPB_REQUIRE(y == nullptr || (scale != nullptr && (!layernorm || nbias != nullptr)), "resid_norm_fwd: norm parameters missing");
  if (M <= 0) return 0;
  if (layernorm)
    resid_norm_fwd_kernel<true><<<(unsigned)M, 128, 0, STREAM(s)>>>((const bf16*)resid, acc, (int)nsplit, slab_stride, bias,
                                                                    (bf16*)xout, scale, nbias, (bf16*)y, mean, rstd, (int)D);
  else
    resid_norm_fwd_kernel<false><<<(unsigned)M, 128, 0, STREAM(s)>>>((const bf16*)resid, acc, (int)nsplit, slab_stride, bias,
                                                                     (bf16*)xout, scale, nbias, (bf16*)y, mean, rstd, (int)D);
  LAPB_LAUNCH_OK("resid_norm_fwd");
  return 0;
}

// plain (scale != NULL) or adaptive (mod != NULL) RMSNorm; row_idx (optional) gathers source rows; dup writes [y|y].
int lapb200_rmsnorm_fwd(const void* x, int64_t ldx, const int64_t* row_idx, const float* scale, const void* mod,
                        int64_t ldmod, int64_t rows_per_sample, void* y, int64_t ldy, int64_t dup, float* rstd,
                        int64_t M, int64_t D, lapb_stream_t s) {
  LAPB_REQUIRE(D % 8 == 0, "rmsnorm: D %% 8 != 0");
  LAPB_REQUIRE((scale != nullptr) != (mod != nullptr), "rmsnorm: exactly one of scale / mod");
  if (scale && !row_idx && !dup && ldx == D && ldy == D && D <= 2048 && M >= 1024) {
    norm_fwd_warp_kernel<false><<<grid_for(M, 8, 148 * 8), 256, 0, STREAM(s)>>>((const bf16*)x, scale, nullptr, (bf16*)y,
                                                                              nullptr, rstd, M, (int)D);
    LAPB_LAUNCH_OK("rmsnorm_fwd");
    return 0;
  }
  rmsnorm_fwd_kernel<<<(unsigned)M, 128, 0, STREAM(s)>>>((const bf16*)x, ldx, (const long*)row_idx, scale,
                                                         (const bf16*)mod, ldmod,
                                                         rows_per_sample > 0 ? (int)rows_per_sample : 1, (bf16*)y, ldy,
                                                         (int)dup, rstd, (int)D);
  LAPB_LAUNCH_OK("rmsnorm_fwd");
  return 0;
}

int lapb200_rmsnorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const int64_t* row_idx,
                        const float* scale, const float* rstd, const void* dres, void* dx, float* dscale, int64_t M,
                        int64_t D, lapb_stream_t s) {
  LAPB_REQUIRE(D % 8 == 0 && D <= 2048, "rmsnorm_bwd: D must be a multiple of 8 and <= 2048");
  LAPB_REQUIRE(lddy % 8 == 0 && ldx % 8 == 0, "rmsnorm_bwd: leading dimensions must be multiples of 8");
  NormBwdArgs a{(const bf16*)dy, lddy, (const bf16*)x, ldx, (const long*)row_idx, scale, nullptr, rstd, (const bf16*)dres,
                (bf16*)dx, dscale, nullptr, M, (int)D};
  if (M <= 0) return 0;
  if (int rc = launch_norm_bwd<false>(a, STREAM(s))) return rc;
  LAPB_LAUNCH_OK("rmsnorm_bwd");
  return 0;
}

int lapb200_ada_rmsnorm_bwd(const void* dy, const void* x, const void* mod, int64_t ldmod, const float* rstd,
                            const void* dres, void* dx, void* dmod, int64_t lddmod, int64_t B,
                            int64_t rows_per_sample, int64_t D, lapb_stream_t s) {
  LAPB_REQUIRE(D % 8 == 0 && D <= 2048, "ada_rmsnorm_bwd: D must be a multiple of 8 and <= 2048");
  ada_rmsnorm_bwd_kernel<<<(unsigned)B, 128, 0, STREAM(s)>>>((const bf16*)dy, (const bf16*)x, (const bf16*)mod, ldmod,
                                                             rstd, (const bf16*)dres, (bf16*)dx, (bf16*)dmod, lddmod,
                                                             (int)rows_per_sample, (int)D);
  LAPB_LAUNCH_OK("ada_rmsnorm_bwd");
  return 0;
}

int lapb200_gated_bwd(const void* dxo, const void* y, const void* gate, int64_t ldg, void* dy, void* dgate,
                      int64_t lddg, int64_t B, int64_t rows_per_sample, int64_t D, lapb_stream_t s) {
  LAPB_REQUIRE(D % 8 == 0, "gated_bwd: D %% 8 != 0");
  gated_bwd_kernel<<<(unsigned)B, 128, 0, STREAM(s)>>>((const bf16*)dxo, (const bf16*)y, (const bf16*)gate, ldg,
                                                       (bf16*)dy, (bf16*)dgate, lddg, (int)rows_per_sample, (int)D);
  LAPB_LAUNCH_OK("gated_bwd");
  return 0;
}

// t_begin = 0 for the joint [prefix|suffix] pass; t_begin = P for a suffix-only pass against a filled cache.
// positions is [B, T - t_begin]; Q is [B, T - t_begin, NH, HD].
int lapb200_rope_fwd(const void* qkv0, const void* qkv1, const int32_t* positions, const float* timescale, void* Q,
                     void* Kc, void* Vc, int64_t B, int64_t P, int64_t A, int64_t Tpad, int64_t NH, int64_t HD,
                     int64_t t_begin, float qscale, lapb_stream_t s) {
  LAPB_REQUIRE(HD % 16 == 0, "rope: head_dim %% 16 != 0");
  long nt = P + A - t_begin;
  LAPB_REQUIRE(nt > 0, "rope: empty token range");
  rope_fwd_kernel<<<(unsigned)(B * nt), 128, 0, STREAM(s)>>>((const bf16*)qkv0, (const bf16*)qkv1, positions,
                                                            timescale, (bf16*)Q, (bf16*)Kc, (bf16*)Vc, (int)P, (int)A,
                                                            (int)Tpad, (int)NH, (int)HD, (int)t_begin, (int)nt,
                                                            qscale);
  LAPB_LAUNCH_OK("rope_fwd");
  return 0;
}

int lapb200_rope_bwd(const void* dQ, const void* dK, const void* dV, const int32_t* positions,
                     const float* timescale, void* dqkv0, void* dqkv1, int64_t B, int64_t P, int64_t A, int64_t Tpad,
                     int64_t NH, int64_t HD, float qscale, lapb_stream_t s) {
  rope_bwd_kernel<<<(unsigned)(B * (P + A)), 128, 0, STREAM(s)>>>((const bf16*)dQ, (const bf16*)dK, (const bf16*)dV,
                                                                  positions, timescale, (bf16*)dqkv0, (bf16*)dqkv1,
                                                                  (int)P, (int)A, (int)Tpad, (int)NH, (int)HD, qscale);
  LAPB_LAUNCH_OK("rope_bwd");
  return 0;
}

int lapb200_geglu_bwd(void* dact, void* gu, int64_t M, int64_t F, int64_t write_act, lapb_stream_t s) {
  LAPB_REQUIRE(F % 8 == 0, "geglu_bwd: F %% 8 != 0");
  if (write_act) geglu_bwd_kernel<true><<<grid_for(M * F / 8, 256), 256, 0, STREAM(s)>>>((bf16*)dact, (bf16*)gu, M, F);
  else geglu_bwd_kernel<false><<<grid_for(M * F / 8, 256), 256, 0, STREAM(s)>>>((bf16*)dact, (bf16*)gu, M, F);
  LAPB_LAUNCH_OK("geglu_bwd");
  return 0;
}

int lapb200_gelu_bwd(void* dh, const void* pre, int64_t n, lapb_stream_t s) {
  LAPB_REQUIRE(n % 8 == 0, "gelu_bwd: n %% 8 != 0");
  gelu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, STREAM(s)>>>((bf16*)dh, (const bf16*)pre, n);
  LAPB_LAUNCH_OK("gelu_bwd");
  return 0;
}

int lapb200_swish_fwd(const float* z, float* y, void* y_bf16, int64_t n, lapb_stream_t s) {
  swish_fwd_kernel<<<grid_for(n, 256), 256, 0, STREAM(s)>>>(z, y, (bf16*)y_bf16, n);
  LAPB_LAUNCH_OK("swish_fwd");
  return 0;
}

int lapb200_swish_bwd(const float* z, const float* dy, const void* dy_bf16, float* dz, int64_t n, lapb_stream_t s) {
  swish_bwd_kernel<<<grid_for(n, 256), 256, 0, STREAM(s)>>>(z, dy, (const bf16*)dy_bf16, dz, n);
  LAPB_LAUNCH_OK("swish_bwd");
  return 0;
}

int lapb200_colsum(const void* X, int64_t ldx, float* out, int64_t M, int64_t N, lapb_stream_t s) {
  LAPB_REQUIRE(N % 2 == 0 && ldx % 2 == 0, "colsum: N, ldx must be even");
  long chunks = 148 * 4 / cdiv(N, 256) + 1;
  long rpc = (M + chunks - 1) / chunks;
  if (rpc < 1) rpc = 1;
  dim3 grid(cdiv(N, 256), cdiv(M, rpc));
  colsum_kernel<<<grid, 128, 0, STREAM(s)>>>((const bf16*)X, ldx, out, M, N, rpc);
  LAPB_LAUNCH_OK("colsum");
  return 0;
}

int lapb200_embed_fwd(const int32_t* ids, const float* E, void* X, int64_t B, int64_t L, int64_t row_off,
                      int64_t rows_per_sample, int64_t D, float scale, lapb_stream_t s) {
  embed_fwd_kernel<<<(unsigned)(B * L), 128, 0, STREAM(s)>>>(ids, E, (bf16*)X, (int)L, row_off, rows_per_sample,
                                                            (int)D, scale);
  LAPB_LAUNCH_OK("embed_fwd");
  return 0;
}

int lapb200_embed_bwd(const int32_t* ids, const void* dX, float* dE, int64_t B, int64_t L, int64_t row_off,
                      int64_t rows_per_sample, int64_t D, float scale, lapb_stream_t s) {
  embed_bwd_kernel<<<(unsigned)(B * L), 128, 0, STREAM(s)>>>(ids, (const bf16*)dX, dE, (int)L, row_off,
                                                            rows_per_sample, (int)D, scale);
  LAPB_LAUNCH_OK("embed_bwd");
  return 0;
}

int lapb200_scatter_rows(const void* d, const int64_t* rows, void* dX, int64_t R, int64_t D, lapb_stream_t s) {
  if (R == 0) return 0;
  scatter_rows_kernel<<<(unsigned)R, 128, 0, STREAM(s)>>>((const bf16*)d, (const long*)rows, (bf16*)dX, (int)D);
  LAPB_LAUNCH_OK("scatter_rows");
  return 0;
}

int lapb200_suffix_inputs(const float* actions, const float* noise, const float* time, float* x_t, float* u_t,
                          float* time_emb, int64_t B, int64_t AD, int64_t W, lapb_stream_t s) {
  suffix_inputs_kernel<<<(unsigned)B, 128, 0, STREAM(s)>>>(actions, noise, time, x_t, u_t, time_emb, (int)B, (int)AD,
                                                           (int)W);
  LAPB_LAUNCH_OK("suffix_inputs");
  return 0;
}

int lapb200_axpy(float* x, const float* v, float dt, int64_t n, lapb_stream_t s) {
  axpy_kernel<<<grid_for(n, 256), 256, 0, STREAM(s)>>>(x, v, dt, n);
  LAPB_LAUNCH_OK("axpy");
  return 0;
}

}  // extern "C"
