// K3 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[b][M,N] = epilogue( A[b][M,K] * B[b][N,K]^T ),  bf16 in, fp32 accumulate in TMEM.
//
// Roles (256 threads, one CTA per SM, static round-robin tile schedule):
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      MMA issuer     (one thread issues tcgen05.mma; tcgen05.commit frees smem slots / publishes TMEM)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue       (tcgen05.ld TMEM -> registers -> fused epilogue -> 128-bit global stores)
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operand layouts: either operand may be K-major (row = M/N index, K contiguous) or MN-major
// (row = K index, M/N contiguous) so forward (K,K), dgrad (K,MN) and wgrad (MN,MN) GEMMs all read the
// tensors where they lie — no transposes are ever materialised.
//
// Reference ops this replaces: OP/models/lora.py:57,145; src/lap/models/backbones/gemma.py:186-201,285;
// OP/models/siglip.py:69-72,88-93,286 and their autodiff transposes.
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"

namespace lapb {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int EPI_THREADS = 256;

struct GemmKArgs {
  int M, N, K;
  int batch_i, batch_o;
  int num_m, num_n, num_k, group_m;
  void* C;
  long ldc, c_bs_i, c_bs_o;
  int c_fp32, accumulate, epi;
  const float* bias;
  const __nv_bfloat16* resid;
  long ldr, r_bs_i, r_bs_o;
  const __nv_bfloat16* gate;
  long ldg;
  int gate_rows;
  __nv_bfloat16* C2;
  long ldc2;
  int q_cols;
  float q_div;
  int a_bi, a_bo, b_bi, b_bo;  // 1 if the operand really has that batch dimension, 0 = broadcast
  int k_splits, kb_per_split;  // split-K (fp32 atomic-add epilogue) for output-starved wgrads
  long split_stride;           // > 0: split s STORES its partial sums to C + s * split_stride (deterministic slabs)
};

template <int BN, bool DUAL, int CG>
struct GemmCfg {
  // CG = CTAs per tile (1, or 2 = a cta_group::2 pair: 256-row tiles, each CTA stages its own 128 rows of A and
  // half of the B tile, one tcgen05.mma M=256 feeds both SMs' tensor cores).
  // DUAL (GeGLU): the MMA tile is N = 2*BN wide: columns [0,BN) come from the gate rows of the stacked [2N,K]
  // weight, columns [BN,2BN) from the up rows, so ONE tcgen05.mma per K step produces both accumulators.
  static constexpr int MMA_N = DUAL ? 2 * BN : BN;
  static constexpr int BN_CTA = MMA_N / CG;          // B rows staged per CTA
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN_CTA * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 8 * 2048;  // per epilogue warp: 32 rows x 64 B transpose buffer
  static constexpr int STAGES_RAW = ((232448 - 1024 - 512 - STAGING_BYTES) / STAGE_BYTES);
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_COLS = MMA_N;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;  // 256 or 512 (power of two)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 512 /*barriers*/;
};

__device__ __forceinline__ void decode_tile(int unit, const GemmKArgs& a, int& bo, int& bi, int& m_blk, int& n_blk,
                                            int& kb0, int& kb1) {
  int tile = unit / a.k_splits;
  int split = unit - tile * a.k_splits;
  kb0 = split * a.kb_per_split;
  kb1 = min(a.num_k, kb0 + a.kb_per_split);
  int per_batch = a.num_m * a.num_n;
  int b = tile / per_batch;
  int t = tile - b * per_batch;
  bo = b / a.batch_i;
  bi = b - bo * a.batch_i;
  int per_group = a.group_m * a.num_n;
  int g = t / per_group;
  int first_m = g * a.group_m;
  int gsize = min(a.num_m - first_m, a.group_m);
  int r = t - g * per_group;
  m_blk = first_m + (r % gsize);
  n_blk = r / gsize;
}

// ---------------------------------------------------------------------------------------------
// fused epilogue on one 8-column vector of one row
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* p, const float (&x)[8]) {
  uint4 v;
  v.x = pack_bf16x2(x[0], x[1]);
  v.y = pack_bf16x2(x[2], x[3]);
  v.z = pack_bf16x2(x[4], x[5]);
  v.w = pack_bf16x2(x[6], x[7]);
  *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ void load_bf16x8(const __nv_bfloat16* p, float (&x)[8]) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = unpack_bf16x2(v.x); x[0] = f.x; x[1] = f.y;
  f = unpack_bf16x2(v.y); x[2] = f.x; x[3] = f.y;
  f = unpack_bf16x2(v.z); x[4] = f.x; x[5] = f.y;
  f = unpack_bf16x2(v.w); x[6] = f.x; x[7] = f.y;
}

template <bool DUAL>
__device__ __forceinline__ void epilogue_vec8(const GemmKArgs& a, long row, int col, long c_boff, long r_boff,
                                              float (&x)[8], float (&x2)[8]) {
  if (DUAL) {
    // GeGLU: g = bf16(acc_g), u = bf16(acc_u); act = bf16( bf16(gelu(g)) * u )
    float act[8], g[8], u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      g[j] = bf16r(x[j]);
      u[j] = bf16r(x2[j]);
      act[j] = bf16r(gelu_tanh(g[j])) * u[j];
    }
    store_bf16x8(reinterpret_cast<__nv_bfloat16*>(a.C) + c_boff + row * a.ldc + col, act);
    if (a.C2) {
      __nv_bfloat16* gu = a.C2 + row * a.ldc2;
      store_bf16x8(gu + col, g);
      store_bf16x8(gu + a.N + col, u);
    }
    return;
  }
  if (a.bias) {
    float4 b0 = *reinterpret_cast<const float4*>(a.bias + col);
    float4 b1 = *reinterpret_cast<const float4*>(a.bias + col + 4);
    float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    if (a.c_fp32) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] += bb[j];
    } else {
      // flax Dense(dtype=bf16): y = bf16(bf16(acc) + bf16(bias))
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = bf16r(x[j]) + bf16r(bb[j]);
    }
  }
  switch (a.epi) {
    case LAPB_EPI_BIAS_GELU: {
      float pre[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        pre[j] = bf16r(x[j]);
        x[j] = gelu_tanh(pre[j]);
      }
      if (a.C2) store_bf16x8(a.C2 + c_boff + row * a.ldc2 + col, pre);
      break;
    }
    case LAPB_EPI_RESID: {
      float r[8];
      load_bf16x8(a.resid + r_boff + row * a.ldr + col, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = r[j] + bf16r(x[j]);
      break;
    }
    case LAPB_EPI_GATED_RESID: {
      float r[8], gt[8];
      load_bf16x8(a.resid + r_boff + row * a.ldr + col, r);
      load_bf16x8(a.gate + (row / a.gate_rows) * a.ldg + col, gt);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = bf16r(x[j]);
      if (a.C2) store_bf16x8(a.C2 + c_boff + row * a.ldc2 + col, x);  // branch output y (needed by dgate)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = r[j] + bf16r(x[j] * gt[j]);
      break;
    }
    case LAPB_EPI_QSCALE: {
      if (col < a.q_cols) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = bf16r(x[j]) * a.q_div;  // q_div holds 1/divisor
      }
      break;
    }
    default:
      break;
  }
  if (a.c_fp32) {
    float* c = reinterpret_cast<float*>(a.C) + c_boff + row * a.ldc + col;
    if (a.k_splits > 1) {  // split-K partial sums: C was zeroed (or holds the value to accumulate onto) by the launcher
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(c + j, x[j]);
      return;
    }
    float4 o0 = make_float4(x[0], x[1], x[2], x[3]);
    float4 o1 = make_float4(x[4], x[5], x[6], x[7]);
    if (a.accumulate) {
      float4 p0 = *reinterpret_cast<float4*>(c);
      float4 p1 = *reinterpret_cast<float4*>(c + 4);
      o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
      o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
    }
    *reinterpret_cast<float4*>(c) = o0;
    *reinterpret_cast<float4*>(c + 4) = o1;
  } else {
    store_bf16x8(reinterpret_cast<__nv_bfloat16*>(a.C) + c_boff + row * a.ldc + col, x);
  }
}

// ---------------------------------------------------------------------------------------------
// coalesced epilogue I/O.  In TMEM a thread owns one ROW of the tile, so naive 16-byte stores from a warp touch 32
// different 128-byte lines per instruction (32 LSU wavefronts) and the epilogue becomes LSU-bound on short-K GEMMs.
// Each epilogue warp therefore transposes its 32x32 bf16 chunk through a private 2 KB smem buffer (16-byte chunks
// XOR-swizzled, conflict-free both ways) so every global instruction moves 8 rows x 64 contiguous bytes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t stg_off(int row, int chunk) {
  return static_cast<uint32_t>(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ void staged_store_bf16(uint8_t* stg, int lane, const uint32_t (&pk)[16],
                                                  __nv_bfloat16* gbase, long ld, int rows_valid, int cols_valid) {
#pragma unroll
  for (int v = 0; v < 4; ++v)
    *reinterpret_cast<uint4*>(stg + stg_off(lane, v)) = make_uint4(pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]);
  __syncwarp();
  const int c = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i;
    if (r < rows_valid && c * 8 < cols_valid)
      *reinterpret_cast<uint4*>(gbase + (long)r * ld + c * 8) = *reinterpret_cast<const uint4*>(stg + stg_off(r, c));
  }
  __syncwarp();
}
// warm L1 with the residual segments a later staged_load_bf16 of the same chunk will read (no registers held)
__device__ __forceinline__ void staged_prefetch(int lane, const __nv_bfloat16* gbase, long ld, int rows_valid,
                                                int cols_valid) {
  const int c = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i;
    if (r < rows_valid && c * 8 < cols_valid)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(gbase + (long)r * ld + c * 8));
  }
}
__device__ __forceinline__ void staged_load_bf16(uint8_t* stg, int lane, const __nv_bfloat16* gbase, long ld,
                                                 int rows_valid, int cols_valid, float (&out)[32]) {
  const int c = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < rows_valid && c * 8 < cols_valid) v = *reinterpret_cast<const uint4*>(gbase + (long)r * ld + c * 8);
    *reinterpret_cast<uint4*>(stg + stg_off(r, c)) = v;
  }
  __syncwarp();
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    uint4 u = *reinterpret_cast<const uint4*>(stg + stg_off(lane, v));
    float2 f;
    f = unpack_bf16x2(u.x); out[8 * v + 0] = f.x; out[8 * v + 1] = f.y;
    f = unpack_bf16x2(u.y); out[8 * v + 2] = f.x; out[8 * v + 3] = f.y;
    f = unpack_bf16x2(u.z); out[8 * v + 4] = f.x; out[8 * v + 5] = f.y;
    f = unpack_bf16x2(u.w); out[8 * v + 6] = f.x; out[8 * v + 7] = f.y;
  }
  __syncwarp();
}

// One 32-row x 32-column chunk of one epilogue warp (bf16 outputs).  row0 = first row of the warp, col0 = first column.
template <bool DUAL>
__device__ __forceinline__ void epilogue_chunk_bf16(const GemmKArgs& a, uint8_t* stg, int lane, long row0, int col0,
                                                    long c_boff, long r_boff, long batch, const uint32_t (&r)[32],
                                                    const uint32_t (&r2)[32]) {
  const int rows_valid = (int)min(32L, (long)a.M - row0);
  const int cols_valid = min(32, a.N - col0);
  if (rows_valid <= 0 || cols_valid <= 0) return;  // warp-uniform
  __nv_bfloat16* C = reinterpret_cast<__nv_bfloat16*>(a.C) + c_boff + row0 * a.ldc + col0;
  uint32_t pk[16];
  if (DUAL) {
    // GeGLU: g = bf16(acc_g), u = bf16(acc_u); act = bf16( bf16(gelu(g)) * u )
    uint32_t pg[16], pu[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float g0 = bf16r(__uint_as_float(r[2 * j])), g1 = bf16r(__uint_as_float(r[2 * j + 1]));
      float u0 = bf16r(__uint_as_float(r2[2 * j])), u1 = bf16r(__uint_as_float(r2[2 * j + 1]));
      pg[j] = pack_bf16x2(g0, g1);
      pu[j] = pack_bf16x2(u0, u1);
      pk[j] = pack_bf16x2(bf16r(gelu_tanh(g0)) * u0, bf16r(gelu_tanh(g1)) * u1);
    }
    staged_store_bf16(stg, lane, pk, C, a.ldc, rows_valid, cols_valid);
    if (a.C2) {
      __nv_bfloat16* gu = a.C2 + row0 * a.ldc2 + col0;
      staged_store_bf16(stg, lane, pg, gu, a.ldc2, rows_valid, cols_valid);
      staged_store_bf16(stg, lane, pu, gu + a.N, a.ldc2, rows_valid, cols_valid);
    }
    return;
  }
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]);
  if (a.bias && a.epi != LAPB_EPI_SOFTMAX_BWD) {
    // flax Dense(dtype=bf16): y = bf16(bf16(acc) + bf16(bias))
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (v * 8 < cols_valid) {
        float4 b0 = *reinterpret_cast<const float4*>(a.bias + col0 + v * 8);
        float4 b1 = *reinterpret_cast<const float4*>(a.bias + col0 + v * 8 + 4);
        float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) x[v * 8 + j] = bf16r(x[v * 8 + j]) + bf16r(bb[j]);
      }
    }
  }
  switch (a.epi) {
    case LAPB_EPI_BIAS_GELU: {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float p0 = bf16r(x[2 * j]), p1 = bf16r(x[2 * j + 1]);
        pk[j] = pack_bf16x2(p0, p1);
        x[2 * j] = gelu_tanh(p0);
        x[2 * j + 1] = gelu_tanh(p1);
      }
      if (a.C2) staged_store_bf16(stg, lane, pk, a.C2 + c_boff + row0 * a.ldc2 + col0, a.ldc2, rows_valid, cols_valid);
      break;
    }
    case LAPB_EPI_RESID: {
      float rr[32];
      staged_load_bf16(stg, lane, a.resid + r_boff + row0 * a.ldr + col0, a.ldr, rows_valid, cols_valid, rr);
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = rr[j] + bf16r(x[j]);
      break;
    }
    case LAPB_EPI_GATED_RESID: {
      float rr[32];
      staged_load_bf16(stg, lane, a.resid + r_boff + row0 * a.ldr + col0, a.ldr, rows_valid, cols_valid, rr);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        x[2 * j] = bf16r(x[2 * j]);
        x[2 * j + 1] = bf16r(x[2 * j + 1]);
        pk[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
      }
      if (a.C2)  // branch output y (needed by dgate)
        staged_store_bf16(stg, lane, pk, a.C2 + c_boff + row0 * a.ldc2 + col0, a.ldc2, rows_valid, cols_valid);
      const long row = row0 + lane;
      if (lane < rows_valid) {
        const __nv_bfloat16* gp = a.gate + (row / a.gate_rows) * a.ldg + col0;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (v * 8 < cols_valid) {
            float gt[8];
            load_bf16x8(gp + v * 8, gt);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[v * 8 + j] = rr[v * 8 + j] + bf16r(x[v * 8 + j] * gt[j]);
          }
        }
      }
      break;
    }
    case LAPB_EPI_QSCALE: {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < a.q_cols) x[j] = bf16r(x[j]) * a.q_div;  // q_div holds 1/divisor
      break;
    }
    case LAPB_EPI_GELU_BWD: {
      float pre[32];
      staged_load_bf16(stg, lane, a.C2 + c_boff + row0 * a.ldc2 + col0, a.ldc2, rows_valid, cols_valid, pre);
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = bf16r(x[j]) * gelu_tanh_grad(pre[j]);
      break;
    }
    case LAPB_EPI_SOFTMAX_BWD: {
      // acc = dP = dO V^T (rounded to bf16 like the stored cotangent); dS = P o (dP - delta[row]), delta = rowsum(dO o O)
      float pp[32];
      staged_load_bf16(stg, lane, a.C2 + c_boff + row0 * a.ldc2 + col0, a.ldc2, rows_valid, cols_valid, pp);
      const float dl = (lane < rows_valid) ? a.bias[batch * a.M + row0 + lane] : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = pp[j] * (bf16r(x[j]) - dl);
      break;
    }
    case LAPB_EPI_GEGLU_BWD: {
      // acc = dAct (rounded to bf16 like the stored cotangent); recompute act, emit dg/du over g/u in place
      __nv_bfloat16* gu = a.C2 + row0 * a.ldc2 + col0;
      float g[32], u[32];
      staged_load_bf16(stg, lane, gu, a.ldc2, rows_valid, cols_valid, g);
      staged_load_bf16(stg, lane, gu + a.N, a.ldc2, rows_valid, cols_valid, u);
      uint32_t pdg[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float d0 = bf16r(x[2 * j]), d1 = bf16r(x[2 * j + 1]);
        float ge0 = bf16r(gelu_tanh(g[2 * j])), ge1 = bf16r(gelu_tanh(g[2 * j + 1]));
        pdg[j] = pack_bf16x2(d0 * u[2 * j] * gelu_tanh_grad(g[2 * j]), d1 * u[2 * j + 1] * gelu_tanh_grad(g[2 * j + 1]));
        pk[j] = pack_bf16x2(d0 * ge0, d1 * ge1);  // du
        x[2 * j] = ge0 * u[2 * j];                 // act
        x[2 * j + 1] = ge1 * u[2 * j + 1];
      }
      staged_store_bf16(stg, lane, pdg, gu, a.ldc2, rows_valid, cols_valid);
      staged_store_bf16(stg, lane, pk, gu + a.N, a.ldc2, rows_valid, cols_valid);
      break;
    }
    default:
      break;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
  staged_store_bf16(stg, lane, pk, C, a.ldc, rows_valid, cols_valid);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, bool DUAL, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmKArgs a) {
  using Cfg = GemmCfg<BN, DUAL, CG>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BN_CTA = Cfg::BN_CTA;
  constexpr int TILE_M = BM * CG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                   // STAGES x 16 KB
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;           // STAGES x B_BYTES
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;     // 8 x 2 KB epilogue transpose buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]  (CG=2: only the leader CTA's copies are used)
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]       (CG=2: leader's copies, armed by both epilogues)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = (CG == 2) ? (int)cluster_ctarank() : 0;
  const int total_tiles = a.num_m * a.num_n * a.batch_i * a.batch_o * a.k_splits;  // work units
  const int first_tile = blockIdx.x / CG;
  const int tile_step = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_THREADS * CG);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    if (CG == 2) { tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch (LAPB_PDL=1): the prologue above touched nothing the preceding kernel writes, so with
  // the launch attribute set this grid is scheduled while its predecessor drains; `wait` blocks until the predecessor has
  // completed and flushed, `launch_dependents` lets the NEXT kernel's launch begin.  Both are no-ops in a normal launch.
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer (every CTA stages its own A rows and its share of B) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        int bo, bi, m_blk, n_blk, kb0, kb1;
        decode_tile(tile, a, bo, bi, m_blk, n_blk, kb0, kb1);
        const int abi = bi * a.a_bi, abo = bo * a.a_bo, bbi = bi * a.b_bi, bbo = bo * a.b_bo;
        const int m0 = m_blk * TILE_M + cta_rank * BM;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES * CG);
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          uint64_t* fb = &full_bar[stage];
#define LAPB_TMA(dst, map, c0, c1, c2, c3)                                   \
  do {                                                                       \
    if (CG == 2) tma_load_4d_2sm(dst, map, fb, c0, c1, c2, c3);               \
    else tma_load_4d(dst, map, fb, c0, c1, c2, c3);                           \
  } while (0)
          if (!A_MN) {
            LAPB_TMA(sa, &tmA, kb * BK, m0, abi, abo);
          } else {
#pragma unroll
            for (int t = 0; t < BM / 64; ++t) LAPB_TMA(sa + t * (64 * BK * 2), &tmA, m0 + t * 64, kb * BK, abi, abo);
          }
          if (DUAL) {
            // gate rows [n_blk*BN, +BN) then up rows [N + n_blk*BN, +BN) form one 2*BN-row K-major B tile
            if (CG == 2) {
              LAPB_TMA(sb, &tmB, kb * BK, n_blk * BN + cta_rank * a.N, bbi, bbo);
            } else {
              LAPB_TMA(sb, &tmB, kb * BK, n_blk * BN, bbi, bbo);
              LAPB_TMA(sb + BN * BK * 2, &tmB, kb * BK, n_blk * BN + a.N, bbi, bbo);
            }
          } else {
            const int n0 = n_blk * BN + cta_rank * BN_CTA;
            if (!B_MN) {
              LAPB_TMA(sb, &tmB, kb * BK, n0, bbi, bbo);
            } else {
#pragma unroll
              for (int t = 0; t < BN_CTA / 64; ++t)
                LAPB_TMA(sb + t * (64 * BK * 2), &tmB, n0 + t * 64, kb * BK, bbi, bbo);
            }
          }
#undef LAPB_TMA
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, Cfg::MMA_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // K-major, SW128: rows 128 B apart, 8-row groups 1024 B apart (SBO); LBO unused (=16).
      // MN-major, SW128: 64-element MN atoms (BK rows x 128 B = 8 KB apart, LBO), 8-row K groups 1024 B (SBO).
      constexpr uint32_t A_LBO = A_MN ? (64 * BK * 2) : 16, A_SBO = 1024;
      constexpr uint32_t B_LBO = B_MN ? (64 * BK * 2) : 16, B_SBO = 1024;
      constexpr uint32_t A_KSTEP = A_MN ? (16 * 128) : 32;  // bytes per UMMA_K=16 step
      constexpr uint32_t B_KSTEP = B_MN ? (16 * 128) : 32;
      // descriptors differ between stages / K steps only in the 14-bit start-address field of the low word:
      // build stage-0 descriptors once, then one 32-bit add per MMA.
      const uint64_t da0 = make_smem_desc_sw128(smem_u32(smem_a), A_LBO, A_SBO);
      const uint64_t db0 = make_smem_desc_sw128(smem_u32(smem_b), B_LBO, B_SBO);
      const uint32_t da_hi = (uint32_t)(da0 >> 32), db_hi = (uint32_t)(db0 >> 32);
      const uint32_t da_lo0 = (uint32_t)da0, db_lo0 = (uint32_t)db0;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        int bo, bi, m_blk, n_blk, kb0, kb1;
        decode_tile(tile, a, bo, bi, m_blk, n_blk, kb0, kb1);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = da_lo0 + stage * (Cfg::A_BYTES >> 4);
          const uint32_t b_lo = db_lo0 + stage * (Cfg::B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = ((uint64_t)da_hi << 32) | (a_lo + k * (A_KSTEP >> 4));
            const uint64_t db = ((uint64_t)db_hi << 32) | (b_lo + k * (B_KSTEP >> 4));
            const uint32_t accum = ((kb - kb0) | k) != 0 ? 1u : 0u;
            if (CG == 2) umma_bf16_2sm(d_tmem, da, db, idesc, accum); else umma_bf16(d_tmem, da, db, idesc, accum);
          }
          // frees this smem slot (in both CTAs) when the MMAs above retire
          if (CG == 2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps of both CTAs
        if (CG == 2) umma_commit_2sm(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // 8 epilogue warps: warp w reads TMEM lane quadrant w % 4 (hardware rule) and column half (w - 4) / 4, so two
    // warps share each SM sub-partition and hide each other's MUFU / conversion latency.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint8_t* stg = staging + (warp - 4) * 2048;
    constexpr int CHUNKS = BN / 32;          // 32-column chunks per accumulator
    constexpr int CH_PER_WARP = CHUNKS / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
      int bo, bi, m_blk, n_blk, kb0, kb1;
      decode_tile(tile, a, bo, bi, m_blk, n_blk, kb0, kb1);
      const long row = (long)m_blk * TILE_M + cta_rank * BM + q * 32 + lane;
      const long c_boff = bi * a.c_bs_i + bo * a.c_bs_o;
      const long r_boff = bi * a.r_bs_i + bo * a.r_bs_o;
      const bool has_resid = !DUAL && !a.c_fp32 && (a.epi == LAPB_EPI_RESID || a.epi == LAPB_EPI_GATED_RESID);
      const long row0w = row - lane;
      const int rows_valid_w = (int)min(32L, (long)a.M - row0w);
      if (has_resid && rows_valid_w > 0) {  // residual rows of this warp's first chunk: fetch while the MMAs still run
        const int col0 = n_blk * BN + half * CH_PER_WARP * 32;
        staged_prefetch(lane, a.resid + r_boff + row0w * a.ldr + col0, a.ldr, rows_valid_w, min(32, a.N - col0));
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::ACC_COLS;
#pragma unroll 1
      for (int cc = 0; cc < CH_PER_WARP; ++cc) {
        const int c = half * CH_PER_WARP + cc;
        uint32_t r[32];
        uint32_t r2[32];
        tmem_ld_32x32(taddr + c * 32, r);
        if (DUAL) tmem_ld_32x32(taddr + BN + c * 32, r2);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (has_resid && rows_valid_w > 0 && cc + 1 < CH_PER_WARP)
          staged_prefetch(lane, a.resid + r_boff + row0w * a.ldr + col0 + 32, a.ldr, rows_valid_w, min(32, a.N - col0 - 32));
        if (!a.c_fp32) {
          epilogue_chunk_bf16<DUAL>(a, stg, lane, row - lane, col0, c_boff, r_boff, (long)bo * a.batch_i + bi, r, r2);
        } else if (a.k_splits > 1) {
          // split-K partial sums: transpose 32x16 fp32 halves through the staging buffer so that each
          // red.global.add.v4.f32 covers 8 rows x 64 contiguous bytes (C was zeroed by the launcher)
          const int rows_valid = (int)min(32L, (long)a.M - row0w);
          if (rows_valid > 0) {
            float* cbase = reinterpret_cast<float*>(a.C) + c_boff + row0w * a.ldc + col0;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
              for (int v = 0; v < 4; ++v)
                *reinterpret_cast<uint4*>(stg + stg_off(lane, v)) =
                    make_uint4(r[hf * 16 + 4 * v], r[hf * 16 + 4 * v + 1], r[hf * 16 + 4 * v + 2], r[hf * 16 + 4 * v + 3]);
              __syncwarp();
              const int cq = lane & 3;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rr = (lane >> 2) + 8 * i;
                const int col = col0 + hf * 16 + cq * 4;
                if (rr < rows_valid && col < a.N) {
                  float4 v4 = *reinterpret_cast<const float4*>(stg + stg_off(rr, cq));
                  float* dst = cbase + (long)rr * a.ldc + hf * 16 + cq * 4;
                  if (a.split_stride > 0)  // deterministic: every split owns a slab, the consumer sums them in order
                    *reinterpret_cast<float4*>(dst + (long)(kb0 / a.kb_per_split) * a.split_stride) = v4;
                  else
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v4.x), "f"(v4.y), "f"(v4.z),
                                 "f"(v4.w)
                                 : "memory");
                }
              }
              __syncwarp();
            }
          }
        } else if (row < a.M) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            int col = col0 + v * 8;
            if (col < a.N) {
              float x[8], x2[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                x[j] = __uint_as_float(r[v * 8 + j]);
                x2[j] = 0.f;
              }
              epilogue_vec8<false>(a, row, col, c_boff, r_boff, x, x2);
            }
          }
        }
      }
      tc_fence_before();
      if (CG == 2) mbar_arrive_cluster(&tempty_bar[acc], 0); else mbar_arrive_relaxed(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Build a rank-4 bf16 tensor map with 128B swizzle. dims/strides innermost first; strides in elements.
int make_tmap_bf16_4d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                      int64_t s1, int64_t s2, int64_t s3, uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = get_encode_fn();
  LAPB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  LAPB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand base must be 16-byte aligned");
  LAPB_REQUIRE(s1 % 8 == 0, "TMA operand leading dimension must be a multiple of 8 elements (got %ld)", (long)s1);
  if (s2 <= 0) s2 = s1 * (int64_t)d1;
  if (s3 <= 0) s3 = s2 * (int64_t)d2;
  LAPB_REQUIRE(s2 % 8 == 0 && s3 % 8 == 0, "TMA batch strides must be multiples of 8 elements");
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {(cuuint64_t)s1 * 2, (cuuint64_t)s2 * 2, (cuuint64_t)s3 * 2};
  cuuint32_t box[4] = {box0, box1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LAPB_REQUIRE(r == CUDA_SUCCESS,
               "cuTensorMapEncodeTiled failed (%d): dims=(%lu,%lu,%lu,%lu) strides=(%ld,%ld,%ld) box=(%u,%u)", (int)r,
               (unsigned long)d0, (unsigned long)d1, (unsigned long)d2, (unsigned long)d3, (long)s1, (long)s2,
               (long)s3, box0, box1);
  return 0;
}

template <int BN, bool A_MN, bool B_MN, bool DUAL, int CG>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKArgs& ka, int grid,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN, DUAL, CG>;
  auto kern = gemm_bf16_tcgen05<BN, A_MN, B_MN, DUAL, CG>;
  static bool configured = false;
  if (!configured) {
    LAPB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  LAPB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, ka));
  return 0;
}

template <int BN, bool DUAL, int CG>
static int dispatch_major(bool amn, bool bmn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKArgs& ka,
                          int grid, cudaStream_t stream) {
  if (DUAL || (!amn && !bmn)) return launch_gemm<BN, false, false, DUAL, CG>(tmA, tmB, ka, grid, stream);
  if (!amn && bmn) return launch_gemm<BN, false, true, false, CG>(tmA, tmB, ka, grid, stream);
  if (amn && bmn) return launch_gemm<BN, true, true, false, CG>(tmA, tmB, ka, grid, stream);
  return launch_gemm<BN, true, false, false, CG>(tmA, tmB, ka, grid, stream);
}

}  // namespace lapb

using namespace lapb;

extern "C" int lapb200_gemm_bf16(const lapb_gemm_t* p, lapb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LAPB_REQUIRE(p != nullptr, "null gemm params");
  LAPB_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0, "gemm: M,N,K must be positive (got %d,%d,%d)", p->M, p->N, p->K);
  LAPB_REQUIRE(p->N % 8 == 0, "gemm: N must be a multiple of 8 (got %d)", p->N);
  LAPB_REQUIRE(p->A && p->B && p->C, "gemm: null operand pointer");
  const bool dual = p->epi == LAPB_EPI_GEGLU;
  LAPB_REQUIRE(!dual || (p->a_major == 0 && p->b_major == 0), "gemm: GEGLU epilogue needs K-major operands");
  LAPB_REQUIRE(!(p->accumulate && !p->c_fp32), "gemm: accumulate requires fp32 C");
  LAPB_REQUIRE(p->ldc % 8 == 0, "gemm: ldc must be a multiple of 8");
  if (p->epi == LAPB_EPI_RESID || p->epi == LAPB_EPI_GATED_RESID)
    LAPB_REQUIRE(p->resid != nullptr && p->ldr % 8 == 0, "gemm: residual epilogue needs resid with ldr%%8==0");
  if (p->epi == LAPB_EPI_GATED_RESID)
    LAPB_REQUIRE(p->gate != nullptr && p->gate_rows > 0 && p->ldg % 8 == 0, "gemm: gated epilogue needs gate");
  if (p->epi == LAPB_EPI_GEGLU_BWD || p->epi == LAPB_EPI_GELU_BWD)
    LAPB_REQUIRE(p->C2 != nullptr && p->ldc2 % 8 == 0 && !p->c_fp32, "gemm: activation-backward epilogue needs bf16 C and C2");
  if (p->epi == LAPB_EPI_SOFTMAX_BWD)
    LAPB_REQUIRE(p->C2 != nullptr && p->ldc2 % 8 == 0 && !p->c_fp32 && p->bias != nullptr,
                 "gemm: softmax-backward epilogue needs bf16 C, P in C2 and the row vector delta in bias");

  int bi = p->batch_i > 0 ? p->batch_i : 1, bo = p->batch_o > 0 ? p->batch_o : 1;
  int BN;
  if (dual) {
    BN = 128;
  } else if (p->block_n == 128 || p->block_n == 256) {
    BN = p->block_n;
  } else {
    // N=256 MMAs run the tensor pipe at twice the work per issue/barrier round trip of N=128 ones: prefer them
    // unless padding N up to a multiple of 256 wastes more than ~20 % of the tile columns.
    long n256 = (long)cdiv(p->N, 256) * 256;
    BN = (p->N > 128 && n256 * 10 <= (long)p->N * 12) ? 256 : 128;
  }
  // 2-CTA pairs (256-row tiles) whenever there are at least two 128-row blocks to pair up
  int CG = (p->cta_group == 1 || p->cta_group == 2) ? p->cta_group : (p->M > 128 ? 2 : 1);
  if (p->cta_group == 0 && p->block_n == 0 && !dual && CG == 2) {
    // small problems (inference prefix pass, expert rows): if 256x256 pair tiles would leave most SMs idle, fall back
    // to 128x128 single-CTA tiles to put more CTAs (and more TMA streams) to work
    long t2 = (long)cdiv(p->M, 256) * cdiv(p->N, BN) * bi * bo * 2;
    long t1 = (long)cdiv(p->M, 128) * cdiv(p->N, 128) * bi * bo;
    if (t2 < num_sms() / 2 && t1 > t2) { CG = 1; BN = 128; }
  }
  const int TILE_M = BM * CG;

  GemmKArgs ka;
  ka.M = p->M; ka.N = p->N; ka.K = p->K;
  ka.batch_i = bi; ka.batch_o = bo;
  ka.num_m = cdiv(p->M, TILE_M);
  ka.num_n = cdiv(p->N, BN);
  ka.num_k = cdiv(p->K, BK);
  ka.group_m = CG == 2 ? 8 : 16;
  ka.C = p->C; ka.ldc = p->ldc; ka.c_bs_i = p->c_bs_i; ka.c_bs_o = p->c_bs_o;
  ka.c_fp32 = p->c_fp32; ka.accumulate = p->accumulate; ka.epi = p->epi;
  ka.bias = p->bias;
  ka.resid = reinterpret_cast<const __nv_bfloat16*>(p->resid);
  ka.ldr = p->ldr; ka.r_bs_i = p->r_bs_i; ka.r_bs_o = p->r_bs_o;
  ka.gate = reinterpret_cast<const __nv_bfloat16*>(p->gate);
  ka.ldg = p->ldg; ka.gate_rows = p->gate_rows > 0 ? p->gate_rows : 1;
  ka.C2 = reinterpret_cast<__nv_bfloat16*>(p->C2); ka.ldc2 = p->ldc2;
  ka.q_cols = p->q_cols; ka.q_div = p->q_div != 0.f ? 1.0f / p->q_div : 1.f;

  // a batch dimension with stride 0 is a broadcast: the tensor map gets extent 1 and the kernel passes coordinate 0
  ka.a_bi = (bi > 1 && p->a_bs_i != 0) ? 1 : 0;
  ka.a_bo = (bo > 1 && p->a_bs_o != 0) ? 1 : 0;
  ka.b_bi = (bi > 1 && p->b_bs_i != 0) ? 1 : 0;
  ka.b_bo = (bo > 1 && p->b_bs_o != 0) ? 1 : 0;
  const int a_di = ka.a_bi ? bi : 1, a_do = ka.a_bo ? bo : 1, b_di = ka.b_bi ? bi : 1, b_do = ka.b_bo ? bo : 1;
  CUtensorMap tmA, tmB;
  int rc;
  if (p->a_major == 0)
    rc = make_tmap_bf16_4d(&tmA, p->A, p->K, p->M, a_di, a_do, p->lda, p->a_bs_i, p->a_bs_o, BK, BM);
  else
    rc = make_tmap_bf16_4d(&tmA, p->A, p->M, p->K, a_di, a_do, p->lda, p->a_bs_i, p->a_bs_o, 64, BK);
  if (rc) return rc;
  const uint64_t b_rows = dual ? 2ull * p->N : (uint64_t)p->N;
  if (p->b_major == 0)
    rc = make_tmap_bf16_4d(&tmB, p->B, p->K, b_rows, b_di, b_do, p->ldb, p->b_bs_i, p->b_bs_o, BK, dual ? BN : BN / CG);
  else
    rc = make_tmap_bf16_4d(&tmB, p->B, p->N, p->K, b_di, b_do, p->ldb, p->b_bs_i, p->b_bs_o, 64, BK);
  if (rc) return rc;

  long total = (long)ka.num_m * ka.num_n * bi * bo;
  int max_ctas = p->max_ctas > 0 ? p->max_ctas : num_sms();
  // split-K: an fp32-out GEMM whose output tiles cannot fill the machine (weight gradients: small M x N, huge K) is
  // cut along K into units that add their partial sums atomically.
  ka.k_splits = 1;
  ka.kb_per_split = ka.num_k;
  ka.split_stride = 0;
  LAPB_REQUIRE(p->split_stride <= 0 || (p->c_fp32 && p->k_splits > 1 && p->epi == LAPB_EPI_NONE && !p->bias && bi * bo == 1),
               "gemm: split_stride (slab split-K) needs fp32 C, k_splits > 1, no epilogue, no batch");
  if (p->c_fp32 && p->epi == LAPB_EPI_NONE && !p->bias && bi * bo == 1 && p->k_splits != 1) {
    const long slots = max_ctas / CG;
    int best = 1;
    if (p->k_splits > 1) {
      best = p->k_splits;
    } else {
      double best_eff = (double)total / (double)(cdiv(total, slots) * slots);
      if (best_eff < 0.80) {
        for (int s2 = 2; s2 <= 16; ++s2) {
          if (ka.num_k / s2 < 16) break;  // keep >= 1024 of K per unit
          double eff = (double)(total * s2) / (double)(cdiv(total * s2, slots) * slots);
          if (eff > best_eff + 0.03) { best_eff = eff; best = s2; }
          if (best_eff >= 0.92) break;
        }
      }
    }
    if (best > 1) {
      ka.kb_per_split = cdiv(ka.num_k, best);
      ka.k_splits = cdiv(ka.num_k, ka.kb_per_split);
      if (p->split_stride > 0)
        LAPB_REQUIRE(ka.k_splits == best && p->split_stride % 4 == 0,
                     "gemm: split_stride needs exactly k_splits non-empty splits (K = %d, k_splits = %d)", p->K, best);
      if (!p->accumulate && p->split_stride <= 0)
        LAPB_CUDA_OK(cudaMemset2DAsync(p->C, (size_t)p->ldc * 4, 0, (size_t)p->N * 4, (size_t)p->M, stream));
      total *= ka.k_splits;
      ka.split_stride = p->split_stride > 0 ? p->split_stride : 0;
    }
  }
  long want = total * CG;
  int grid = (int)(want < max_ctas ? want : max_ctas);
  grid -= grid % CG;
  if (grid < CG) grid = CG;

  const bool amn = p->a_major != 0, bmn = p->b_major != 0;
  if (CG == 2) {
    if (dual) return dispatch_major<128, true, 2>(amn, bmn, tmA, tmB, ka, grid, stream);
    if (BN == 256) return dispatch_major<256, false, 2>(amn, bmn, tmA, tmB, ka, grid, stream);
    return dispatch_major<128, false, 2>(amn, bmn, tmA, tmB, ka, grid, stream);
  }
  if (dual) return dispatch_major<128, true, 1>(amn, bmn, tmA, tmB, ka, grid, stream);
  if (BN == 256) return dispatch_major<256, false, 1>(amn, bmn, tmA, tmB, ka, grid, stream);
  return dispatch_major<128, false, 1>(amn, bmn, tmA, tmB, ka, grid, stream);
}
