// K1 — fused Gemma shared-attention forward for sm_100a (gemma.py:234-272), head_dim 256, one KV head.
//
//   S = Q K^T (tcgen05, accumulators in TMEM)  ->  where(mask, S, -2.3819763e38) -> softmax fp32 -> P bf16
//   O = P V   (tcgen05, P staged in shared memory as the A operand, V read MN-major straight from the cache layout)
//
// The 8 query heads of a token share the KV head, so they are stacked into the MMA M dimension: a CTA owns 128
// query rows (16 tokens x 8 heads) of one sample and walks the keys in tiles of 64, K/V tiles double-buffered in
// shared memory by TMA.  The softmax is TWO-PASS: pass 1 only accumulates the row max / sum (online), pass 2
// recomputes S and forms p = exp(s - max) / sum rounded to bf16 — exactly the reference's rounding point — so the
// result is bit-compatible with the unfused (GEMM + softmax + GEMM) path, and P can be written out for the backward
// pass with a TMA store straight from the swizzled A-operand tile.  Attention is 2 % of the model FLOPs, so spending
// a second QK^T to avoid rescaling the 256-column O accumulator in TMEM is the cheap choice.
//
// Warp roles (384 threads): warps 0-7 softmax/epilogue — two warpgroups, thread == query row == TMEM lane, warpgroup g
// owns key columns [32g, 32g+32) of every 64-key tile so two warps share each scheduler and hide each other's
// MUFU/ALU latency; warp 8 TMA producer, warp 9 MMA issuer, warp 10 TMEM allocator.
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"

namespace lapb {

typedef __nv_bfloat16 bf16;
#define BIG_NEG (-2.3819763e38f)

constexpr int FA_QT = 128;  // query rows per CTA
constexpr int FA_KT = 64;   // keys per tile
constexpr int FA_HD = 256;  // head dim
constexpr int FA_Q_BYTES = FA_QT * FA_HD * 2;      // 64 KB: 4 k-chunks of [128 rows x 128 B]
constexpr int FA_KV_BYTES = FA_KT * FA_HD * 2;     // 32 KB per stage
constexpr int FA_P_BYTES = FA_QT * FA_KT * 2;      // 16 KB
constexpr int FA_SMEM = FA_Q_BYTES + 4 * FA_KV_BYTES + FA_P_BYTES + 1024 + 256 + 2048;  // + (m,l) exchange

struct FaArgs {
  int B, R, G, Tq, S_len, Tpad, W32, NT;
  const uint32_t* bits;
  bf16* O0;
  bf16* O1;
  int split_row;  // rows [0, split_row) of a sample -> O0, the rest -> O1
  int write_p;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(384, 1)
fa_gemma_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP, const FaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;
  uint8_t* Ks = smem + FA_Q_BYTES;
  uint8_t* Vs = Ks + 2 * FA_KV_BYTES;
  uint8_t* Ps = Vs + 2 * FA_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ps + FA_P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* p_empty = bars + 14;
  uint64_t* o_full = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * FA_QT;
  const int NT = a.NT;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 9 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 256);
    }
    mbar_init(p_full, 1);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 10) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 128;  // S buffers: columns [0,64) and [64,128); O: [128, 384)

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, FA_Q_BYTES);
#pragma unroll
      for (int c = 0; c < 4; ++c) tma_load_4d(Qs + c * (FA_QT * 128), &tmQ, q_full, c * 64, q0, b, 0);
      int it = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < NT; ++i, ++it) {
          const int st = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_expect_tx(&k_full[st], FA_KV_BYTES);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            tma_load_4d(Ks + st * FA_KV_BYTES + c * (FA_KT * 128), &tmK, &k_full[st], c * 64, i * FA_KT, b, 0);
          if (pass == 1) {
            const int vst = i & 1;
            const uint32_t vph = (i >> 1) & 1;
            mbar_wait(&v_empty[vst], vph ^ 1);
            mbar_expect_tx(&v_full[vst], FA_KV_BYTES);
#pragma unroll
            for (int c = 0; c < 4; ++c)  // 64-dim atoms of the MN-major B operand: [64 keys x 128 B] each
              tma_load_4d(Vs + vst * FA_KV_BYTES + c * (FA_KT * 128), &tmV, &v_full[vst], c * 64, i * FA_KT, b, 0);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idescS = make_idesc_bf16(FA_QT, FA_KT, 0, 0);
      constexpr uint32_t idescPV = make_idesc_bf16(FA_QT, FA_HD, 0, 1);
      const uint32_t q_addr = smem_u32(Qs), k_addr = smem_u32(Ks), v_addr = smem_u32(Vs), p_addr = smem_u32(Ps);
      mbar_wait(q_full, 0);
      auto issue_S = [&](int it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&k_full[st], ph);
        mbar_wait(&s_empty[st], ph ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + st * FA_KT;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint64_t da = make_smem_desc_sw128(q_addr + c * (FA_QT * 128) + kk * 32, 16, 1024);
            uint64_t db = make_smem_desc_sw128(k_addr + st * FA_KV_BYTES + c * (FA_KT * 128) + kk * 32, 16, 1024);
            umma_bf16(d, da, db, idescS, (c | kk) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[st]);
      };
      for (int it = 0; it < NT; ++it) issue_S(it);  // pass 1: logits only
      issue_S(NT);                                   // pass 2, software-pipelined: S(i+1) is issued before PV(i)
      for (int i = 0; i < NT; ++i) {
        if (i + 1 < NT) issue_S(NT + i + 1);
        const int vst = i & 1;
        mbar_wait(p_full, i & 1);
        mbar_wait(&v_full[vst], (i >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint64_t da = make_smem_desc_sw128(p_addr + kk * 32, 16, 1024);
          // V tile: MN-major, 4 atoms of 64 dims ([64 keys x 128 B] = 8 KB apart), 16 keys per step = 2 KB
          uint64_t db = make_smem_desc_sw128(v_addr + vst * FA_KV_BYTES + kk * (16 * 128), FA_KT * 128, 1024);
          umma_bf16(tmem_O, da, db, idescPV, (i | kk) != 0 ? 1u : 0u);
        }
        umma_commit(p_empty);
        umma_commit(&v_empty[vst]);
      }
      umma_commit(o_full);
    }
  } else if (warp < 8) {
    // ===================== softmax + epilogue: thread == (query row, 32-key column half) =====================
    const int wg = warp >> 2;                 // column half of every key tile / dim half of the output
    const int r = (warp & 3) * 32 + lane;     // query row of the tile == TMEM lane
    const long grow = (long)q0 + r;
    const bool valid_row = grow < a.R;
    long tok = grow / a.G;
    if (tok > a.Tq - 1) tok = a.Tq - 1;
    const uint32_t* mrow = a.bits + ((long)b * a.Tq + tok) * a.W32;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    float* stat = reinterpret_cast<float*>(tmem_slot + 4);  // [2][128][2] (m, l) exchange between the warpgroups
    float m = -3.4e38f, l = 0.f;
    // ---- pass 1: running max / sum over this warpgroup's columns ----
    for (int it = 0; it < NT; ++it) {
      const int sb = it & 1;
      mbar_wait(&s_full[sb], (it >> 1) & 1);
      tc_fence_after();
      uint32_t sv[32];
      tmem_ld_32x32(tmem_base + lane_base + sb * FA_KT + wg * 32, sv);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_relaxed(&s_empty[sb]);
      const int key0 = it * FA_KT + wg * 32;
      const uint32_t w = mrow[2 * it + wg];
      const int nvalid = a.S_len - key0;  // columns [0, nvalid) of this half are real keys
      float tmax = -3.4e38f;
      if (w != 0xFFFFFFFFu) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (!((w >> j) & 1u)) sv[j] = __float_as_uint(BIG_NEG);
      }
      if (nvalid >= 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) tmax = fmaxf(tmax, __uint_as_float(sv[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) tmax = fmaxf(tmax, __uint_as_float(sv[j]));
      }
      const float m_new = fmaxf(m, tmax);
      float sum = 0.f;
      // (x - m) is formed BEFORE scaling by log2(e): for a fully masked row m = -2.38e38 and m*log2(e) would overflow
      if (nvalid >= 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) sum += exp2f((__uint_as_float(sv[j]) - m_new) * LOG2E);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) sum += exp2f((__uint_as_float(sv[j]) - m_new) * LOG2E);
      }
      l = l * exp2f((m - m_new) * LOG2E) + sum;
      m = m_new;
    }
    // combine the two column halves
    stat[(wg * 128 + r) * 2 + 0] = m;
    stat[(wg * 128 + r) * 2 + 1] = l;
    softmax_bar();
    {
      const float mo = stat[((wg ^ 1) * 128 + r) * 2 + 0], lo = stat[((wg ^ 1) * 128 + r) * 2 + 1];
      const float mf = fmaxf(m, mo);
      l = l * exp2f((m - mf) * LOG2E) + lo * exp2f((mo - mf) * LOG2E);
      m = mf;
    }
    // ---- pass 2: p = exp(s - max) / sum -> bf16 -> smem (A operand of P V) [+ TMA store for the backward] ----
    const float inv = 1.0f / l;
    for (int i = 0; i < NT; ++i) {
      const int it = NT + i;
      const int sb = it & 1;
      mbar_wait(&s_full[sb], (it >> 1) & 1);
      tc_fence_after();
      uint32_t sv[32];
      tmem_ld_32x32(tmem_base + lane_base + sb * FA_KT + wg * 32, sv);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_relaxed(&s_empty[sb]);
      const int key0 = i * FA_KT + wg * 32;
      const uint32_t w = mrow[2 * i + wg];
      const int nvalid = a.S_len - key0;
      if (w != 0xFFFFFFFFu) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (!((w >> j) & 1u)) sv[j] = __float_as_uint(BIG_NEG);
      }
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float p0 = exp2f((__uint_as_float(sv[j]) - m) * LOG2E) * inv;
        float p1 = exp2f((__uint_as_float(sv[j + 1]) - m) * LOG2E) * inv;
        if (j >= nvalid) p0 = 0.f;
        if (j + 1 >= nvalid) p1 = 0.f;
        pk[j >> 1] = pack_bf16x2(p0, p1);
      }
      // the P tile may be overwritten once the previous P V MMAs and the previous TMA store have read it
      if (threadIdx.x == 0) {
        mbar_wait(p_empty, (i & 1) ^ 1);
        if (a.write_p) tma_store_wait_read();
      }
      softmax_bar();
      // K-major, 128B-swizzled A tile: row r is 128 B (64 keys); 16-byte chunk c sits at chunk position c ^ (r & 7)
      uint8_t* prow = Ps + r * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int cc = wg * 4 + c;
        *reinterpret_cast<uint4*>(prow + ((cc ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
      fence_proxy_async();
      softmax_bar();
      if (threadIdx.x == 0) {
        if (a.write_p) {
          tma_store_4d(&tmP, Ps, i * FA_KT, q0, b, 0);
          tma_store_commit();
        }
        mbar_arrive(p_full);
      }
    }
    // ---- epilogue: O (fp32, TMEM) -> bf16 rows; warpgroup g stores dims [128g, 128g + 128) ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    bf16* orow = nullptr;
    if (valid_row) {
      if (grow < a.split_row) orow = a.O0 + ((long)b * a.split_row + grow) * FA_HD;
      else orow = a.O1 + ((long)b * (a.R - a.split_row) + (grow - a.split_row)) * FA_HD;
    }
#pragma unroll 1
    for (int c4 = 0; c4 < FA_HD / 64; ++c4) {
      const int c = wg * (FA_HD / 64) + c4;
      uint32_t o[32];
      tmem_ld_32x32(tmem_O + lane_base + c * 32, o);
      tmem_ld_wait();
      if (valid_row) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * v + 0]), __uint_as_float(o[8 * v + 1]));
          u.y = pack_bf16x2(__uint_as_float(o[8 * v + 2]), __uint_as_float(o[8 * v + 3]));
          u.z = pack_bf16x2(__uint_as_float(o[8 * v + 4]), __uint_as_float(o[8 * v + 5]));
          u.w = pack_bf16x2(__uint_as_float(o[8 * v + 6]), __uint_as_float(o[8 * v + 7]));
          *reinterpret_cast<uint4*>(orow + c * 32 + v * 8) = u;
        }
      }
    }
    if (threadIdx.x == 0 && a.write_p) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int make_tmap_bf16_4d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, int64_t s1,
                      int64_t s2, int64_t s3, uint32_t box0, uint32_t box1);  // gemm.cu

}  // namespace lapb

using namespace lapb;

extern "C" int lapb200_fa_gemma_fwd(const void* Q, const void* Kc, const void* Vc, const uint32_t* bits, void* P,
                                    void* O0, void* O1, int64_t B, int64_t R, int64_t G, int64_t Tq, int64_t S_len,
                                    int64_t Tpad, int64_t W32, int64_t split_row, int64_t head_dim,
                                    lapb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LAPB_REQUIRE(head_dim == FA_HD, "fa_gemma_fwd: head_dim must be %d (got %ld)", FA_HD, (long)head_dim);
  LAPB_REQUIRE(Tpad % FA_KT == 0 && S_len <= Tpad && W32 * 32 >= Tpad, "fa_gemma_fwd: Tpad must be a multiple of %d", FA_KT);
  LAPB_REQUIRE(R == Tq * G && split_row >= 0 && split_row <= R, "fa_gemma_fwd: inconsistent row counts");
  CUtensorMap tmQ, tmK, tmV, tmP;
  int rc;
  if ((rc = make_tmap_bf16_4d(&tmQ, Q, FA_HD, R, B, 1, FA_HD, R * FA_HD, 0, 64, FA_QT))) return rc;
  if ((rc = make_tmap_bf16_4d(&tmK, Kc, FA_HD, Tpad, B, 1, FA_HD, Tpad * FA_HD, 0, 64, FA_KT))) return rc;
  if ((rc = make_tmap_bf16_4d(&tmV, Vc, FA_HD, Tpad, B, 1, FA_HD, Tpad * FA_HD, 0, 64, FA_KT))) return rc;
  if (P) {
    if ((rc = make_tmap_bf16_4d(&tmP, P, Tpad, R, B, 1, Tpad, R * Tpad, 0, 64, FA_QT))) return rc;
  } else {
    tmP = tmQ;
  }
  FaArgs a;
  a.B = (int)B; a.R = (int)R; a.G = (int)G; a.Tq = (int)Tq; a.S_len = (int)S_len; a.Tpad = (int)Tpad;
  a.W32 = (int)W32; a.NT = (int)(Tpad / FA_KT);
  a.bits = bits; a.O0 = (bf16*)O0; a.O1 = (bf16*)O1; a.split_row = (int)split_row; a.write_p = P ? 1 : 0;
  static bool configured = false;
  if (!configured) {
    LAPB_CUDA_OK(cudaFuncSetAttribute(fa_gemma_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    configured = true;
  }
  dim3 grid(cdiv(R, FA_QT), (unsigned)B);
  fa_gemma_fwd_kernel<<<grid, 384, FA_SMEM, stream>>>(tmQ, tmK, tmV, tmP, a);
  LAPB_LAUNCH_OK("fa_gemma_fwd");
  return 0;
}
