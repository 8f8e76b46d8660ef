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
constexpr int GEMM_THREADS = 256;

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
};

template <int BN, bool DUAL>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int NB = DUAL ? 2 : 1;
  static constexpr int STAGE_BYTES = A_BYTES + NB * B_BYTES;
  static constexpr int STAGES = (196608 / STAGE_BYTES);  // 4 x 48KB or 6 x 32KB
  static constexpr int ACC_COLS = BN * NB;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;  // 256 or 512 (power of two)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/;
};

__device__ __forceinline__ void decode_tile(int tile, const GemmKArgs& a, int& bo, int& bi, int& m_blk, int& n_blk) {
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
        for (int j = 0; j < 8; ++j) x[j] = bf16r(x[j]) / a.q_div;
      }
      break;
    }
    default:
      break;
  }
  if (a.c_fp32) {
    float* c = reinterpret_cast<float*>(a.C) + c_boff + row * a.ldc + col;
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
// the kernel
// ---------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, bool DUAL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmKArgs a) {
  using Cfg = GemmCfg<BN, DUAL>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                   // STAGES x 16 KB
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;           // STAGES x NB x B_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = a.num_m * a.num_n * a.batch_i * a.batch_o;

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
      mbar_init(&tempty_bar[s], 128);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int bo, bi, m_blk, n_blk;
        decode_tile(tile, a, bo, bi, m_blk, n_blk);
        for (int kb = 0; kb < a.num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * (Cfg::NB * Cfg::B_BYTES);
          const int abi = bi * a.a_bi, abo = bo * a.a_bo, bbi = bi * a.b_bi, bbo = bo * a.b_bo;
          if (!A_MN) {
            tma_load_4d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM, abi, abo);
          } else {
#pragma unroll
            for (int t = 0; t < BM / 64; ++t)
              tma_load_4d(sa + t * (64 * BK * 2), &tmA, &full_bar[stage], m_blk * BM + t * 64, kb * BK, abi, abo);
          }
#pragma unroll
          for (int d = 0; d < Cfg::NB; ++d) {
            // dual: second B tile lives N rows further down the same [2N, K] weight
            int n0 = n_blk * BN + d * a.N;
            uint8_t* sbd = sb + d * Cfg::B_BYTES;
            if (!B_MN) {
              tma_load_4d(sbd, &tmB, &full_bar[stage], kb * BK, n0, bbi, bbo);
            } else {
#pragma unroll
              for (int t = 0; t < BN / 64; ++t)
                tma_load_4d(sbd + t * (64 * BK * 2), &tmB, &full_bar[stage], n0 + t * 64, kb * BK, bbi, bbo);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // K-major, SW128: rows 128 B apart, 8-row groups 1024 B apart (SBO); LBO unused (=16).
      // MN-major, SW128: 64-element MN atoms (BK rows x 128 B = 8 KB apart, LBO), 8-row K groups 1024 B (SBO).
      constexpr uint32_t A_LBO = A_MN ? (64 * BK * 2) : 16, A_SBO = 1024;
      constexpr uint32_t B_LBO = B_MN ? (64 * BK * 2) : 16, B_SBO = 1024;
      constexpr uint32_t A_KSTEP = A_MN ? (16 * 128) : 32;  // bytes per UMMA_K=16 step
      constexpr uint32_t B_KSTEP = B_MN ? (16 * 128) : 32;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        for (int kb = 0; kb < a.num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * (Cfg::NB * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            uint64_t da = make_smem_desc_sw128(sa + k * A_KSTEP, A_LBO, A_SBO);
            uint64_t db = make_smem_desc_sw128(sb + k * B_KSTEP, B_LBO, B_SBO);
            umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            if (DUAL) {
              uint64_t db2 = make_smem_desc_sw128(sb + Cfg::B_BYTES + k * B_KSTEP, B_LBO, B_SBO);
              umma_bf16(d_tmem + BN, da, db2, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);  // frees this smem slot when the MMAs above retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp - 4;  // TMEM lane quadrant == warp % 4
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int bo, bi, m_blk, n_blk;
      decode_tile(tile, a, bo, bi, m_blk, n_blk);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const long row = (long)m_blk * BM + q * 32 + lane;
      const long c_boff = bi * a.c_bs_i + bo * a.c_bs_o;
      const long r_boff = bi * a.r_bs_i + bo * a.r_bs_o;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::ACC_COLS;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        uint32_t r2[32];
        tmem_ld_32x32(taddr + c * 32, r);
        if (DUAL) tmem_ld_32x32(taddr + BN + c * 32, r2);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row < a.M) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            int col = col0 + v * 8;
            if (col < a.N) {
              float x[8], x2[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                x[j] = __uint_as_float(r[v * 8 + j]);
                x2[j] = DUAL ? __uint_as_float(r2[v * 8 + j]) : 0.f;
              }
              epilogue_vec8<DUAL>(a, row, col, c_boff, r_boff, x, x2);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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

template <int BN, bool A_MN, bool B_MN, bool DUAL>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKArgs& ka, int grid,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN, DUAL>;
  auto kern = gemm_bf16_tcgen05<BN, A_MN, B_MN, DUAL>;
  static bool configured = false;
  if (!configured) {
    LAPB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, ka);
  LAPB_LAUNCH_OK("gemm_bf16_tcgen05");
  return 0;
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

  int bi = p->batch_i > 0 ? p->batch_i : 1, bo = p->batch_o > 0 ? p->batch_o : 1;
  int BN;
  if (dual) {
    BN = 128;
  } else if (p->block_n == 128 || p->block_n == 256) {
    BN = p->block_n;
  } else {
    long n128 = (long)cdiv(p->N, 128) * 128, n256 = (long)cdiv(p->N, 256) * 256;
    BN = (p->N >= 256 && n256 == n128) ? 256 : 128;
  }

  GemmKArgs ka;
  ka.M = p->M; ka.N = p->N; ka.K = p->K;
  ka.batch_i = bi; ka.batch_o = bo;
  ka.num_m = cdiv(p->M, BM);
  ka.num_n = cdiv(p->N, BN);
  ka.num_k = cdiv(p->K, BK);
  ka.group_m = 16;
  ka.C = p->C; ka.ldc = p->ldc; ka.c_bs_i = p->c_bs_i; ka.c_bs_o = p->c_bs_o;
  ka.c_fp32 = p->c_fp32; ka.accumulate = p->accumulate; ka.epi = p->epi;
  ka.bias = p->bias;
  ka.resid = reinterpret_cast<const __nv_bfloat16*>(p->resid);
  ka.ldr = p->ldr; ka.r_bs_i = p->r_bs_i; ka.r_bs_o = p->r_bs_o;
  ka.gate = reinterpret_cast<const __nv_bfloat16*>(p->gate);
  ka.ldg = p->ldg; ka.gate_rows = p->gate_rows > 0 ? p->gate_rows : 1;
  ka.C2 = reinterpret_cast<__nv_bfloat16*>(p->C2); ka.ldc2 = p->ldc2;
  ka.q_cols = p->q_cols; ka.q_div = p->q_div != 0.f ? p->q_div : 1.f;

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
    rc = make_tmap_bf16_4d(&tmB, p->B, p->K, b_rows, b_di, b_do, p->ldb, p->b_bs_i, p->b_bs_o, BK, BN);
  else
    rc = make_tmap_bf16_4d(&tmB, p->B, p->N, p->K, b_di, b_do, p->ldb, p->b_bs_i, p->b_bs_o, 64, BK);
  if (rc) return rc;

  long total = (long)ka.num_m * ka.num_n * bi * bo;
  int max_ctas = p->max_ctas > 0 ? p->max_ctas : num_sms();
  int grid = (int)(total < max_ctas ? total : max_ctas);

  const bool amn = p->a_major != 0, bmn = p->b_major != 0;
  if (dual) return launch_gemm<128, false, false, true>(tmA, tmB, ka, grid, stream);
  if (BN == 256) {
    if (!amn && !bmn) return launch_gemm<256, false, false, false>(tmA, tmB, ka, grid, stream);
    if (!amn && bmn) return launch_gemm<256, false, true, false>(tmA, tmB, ka, grid, stream);
    if (amn && bmn) return launch_gemm<256, true, true, false>(tmA, tmB, ka, grid, stream);
    return launch_gemm<256, true, false, false>(tmA, tmB, ka, grid, stream);
  } else {
    if (!amn && !bmn) return launch_gemm<128, false, false, false>(tmA, tmB, ka, grid, stream);
    if (!amn && bmn) return launch_gemm<128, false, true, false>(tmA, tmB, ka, grid, stream);
    if (amn && bmn) return launch_gemm<128, true, true, false>(tmA, tmB, ka, grid, stream);
    return launch_gemm<128, true, false, false>(tmA, tmB, ka, grid, stream);
  }
}
