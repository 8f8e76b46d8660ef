// K1p — EXPERIMENTAL cta_group::2 variant of the fused Gemma attention forward (fa_gemma.cu); selected with LAPB_FA_PAIR=1.
//
// Why: profiles/r01_fa_knobs.md — every 128-row CTA of K1 streams the sample's whole K twice and V once from L2, 64 B/clk/SM
// in pass 1 against the ~42.6 B/clk/SM the L2 delivers to 148 SMs; more than half of K1's time is that floor.  Here two CTAs
// of a cluster form ONE tcgen05 cta_group::2 tile of 256 query rows: each CTA keeps its own 128 rows of Q / S / P / O, and the
// B operand of every MMA is SPLIT across the pair — a CTA loads only half of each K slice (128 of the 256 keys) and half of
// each V tile (128 of the 256 dims).  L2 traffic per FLOP halves; the 16 KB stages leave room for an 8-deep ring.
//
// Same arithmetic, rounding points and outputs as K1 (two-pass softmax, p rounded to bf16, optional P store).  Structure per
// CTA is K1's (4 softmax warpgroups, TMA warp, MMA warp, TMEM warp); what changes:
//   * the LEADER CTA (cluster rank 0) alone issues tcgen05.mma.cta_group::2 / tcgen05.commit (multicast to both CTAs);
//   * barriers the leader's MMA warp waits on live in the leader: q_full and r_full collect the TMA bytes of BOTH CTAs
//     (cp.async.bulk.tensor ... cta_group::2), s_empty counts the softmax threads of both CTAs and p_full both warpgroup
//     leaders (remote mbarrier arrives); r_empty / s_full / pv_done / o_full are per-CTA copies armed by the multicast commit.
// Status: written against the measured analysis at the end of round 1; see DESIGN.md §7 for what has been run on hardware.
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace lapb {

typedef __nv_bfloat16 bf16;
#define BIG_NEG (-2.3819763e38f)

namespace pair {
constexpr int QT = 128;                      // query rows per CTA (256 per pair)
constexpr int KT = 64;                       // keys per P V step
constexpr int KC = 256;                      // keys per S chunk
constexpr int HD = 256;
constexpr int WG = 4;
constexpr int SOFT = 128 * WG;
constexpr int THREADS = SOFT + 96;
constexpr int NST = 8;                       // ring stages of 16 KB
constexpr int NPB = 2;                       // P buffers
constexpr int Q_BYTES = QT * HD * 2;         // 64 KB
constexpr int ST_BYTES = 16 * 1024;          // half a K slice [128 keys x 64 dims] or half a V tile [64 keys x 128 dims]
constexpr int P_BYTES = QT * KT * 2;         // 16 KB
constexpr int SMEM = Q_BYTES + NST * ST_BYTES + NPB * P_BYTES + 1024 + 512;
static_assert(SMEM <= 227 * 1024, "fa_gemma_pair: shared memory");
}  // namespace pair

struct FaPairArgs {
  int B, R, G, Tq, S_len, Tpad, W32, NCH;
  const uint32_t* bits;
  bf16* O0;
  bf16* O1;
  int split_row;
  int write_p;
};

__device__ __forceinline__ void p_tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void p_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void p_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void p_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void p_softmax_bar() { asm volatile("bar.sync 1, %0;" ::"n"(pair::SOFT) : "memory"); }
__device__ __forceinline__ void p_wg_bar(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(2 + wg) : "memory"); }
__device__ __forceinline__ float p_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// arrive with RELEASE semantics at cluster scope on the barrier at this offset in CTA `cta`: publishes this CTA's
// shared-memory writes (P sub-tile, already fenced into the async proxy) to the leader that will issue the MMA reading them
__device__ __forceinline__ void mbar_arrive_cluster_release(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(pair::THREADS, 1)
fa_gemma_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP,
                         const FaPairArgs a) {
  using namespace pair;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;
  uint8_t* Ring = smem + Q_BYTES;
  uint8_t* Ps = Ring + NST * ST_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ps + NPB * P_BYTES);
  uint64_t* q_full = bars;          // leader: Q of both CTAs landed
  uint64_t* r_full = bars + 1;      // [8] leader: both halves of the stage landed
  uint64_t* r_empty = bars + 9;     // [8] per CTA: the MMAs reading the stage retired
  uint64_t* s_full = bars + 17;     // [2] per CTA
  uint64_t* s_empty = bars + 19;    // [2] leader: every softmax thread of both CTAs has pulled S
  uint64_t* p_full = bars + 21;     // [4] leader: sub-tile w written in both CTAs
  uint64_t* pv_done = bars + 25;    // [4] per CTA
  uint64_t* st_done = bars + 29;    // [4] per CTA (own TMA store has read the P buffer)
  uint64_t* o_full = bars + 33;     // per CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 34);
  float* stat = reinterpret_cast<float*>(Ps);  // (m, l) exchange borrows the P buffers (not written before pass 2)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * QT;  // this CTA's rows (the pair covers [q0 & ~255, +256))
  const int NCH = a.NCH;
  constexpr int W_TMA = 4 * WG, W_MMA = W_TMA + 1, W_ALLOC = W_TMA + 2;
  const int U0 = (NCH + 1) / 2;
  auto chunk_keys = [&](int j) { return min(KC, a.Tpad - j * KC); };  // multiple of 64

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == W_MMA && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], 1);
    }
    for (int i = 0; i < WG; ++i) {
      mbar_init(&p_full[i], 2);
      mbar_init(&pv_done[i], 1);
      mbar_init(&st_done[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 2 * SOFT);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == W_ALLOC) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;

  if (warp == W_TMA) {
    // ===================== TMA producer (each CTA: its own Q rows, its half of every K slice / V tile) =====================
    if (lane == 0) {
      if (rank == 0) mbar_expect_tx(q_full, 2 * Q_BYTES);
#pragma unroll
      for (int c = 0; c < 4; ++c) tma_load_4d_2sm(Qs + c * (QT * 128), &tmQ, q_full, c * 64, q0, b, 0);
      int it = 0;
      auto load_k_chunk = [&](int j) {  // four 64-dim slices; this CTA's keys are [j*256 + rank*n/2, +n/2), n = keys of the chunk
        const int half = chunk_keys(j) / 2;
        for (int c = 0; c < 4; ++c, ++it) {
          const int st = it % NST;
          mbar_wait(&r_empty[st], ((it / NST) & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(&r_full[st], 2 * ST_BYTES);
          // the box is always 128 keys: for a short last chunk the surplus rows are loaded and ignored (N/2 rows are used)
          tma_load_4d_2sm(Ring + st * ST_BYTES, &tmK, &r_full[st], c * 64, j * KC + (int)rank * half, b, 0);
        }
      };
      auto load_v_tile = [&](int key0) {  // this CTA's 128 dims: 2 atoms of 64 dims ([64 keys x 128 B] each)
        const int st = it % NST;
        mbar_wait(&r_empty[st], ((it / NST) & 1) ^ 1);
        if (rank == 0) mbar_expect_tx(&r_full[st], 2 * ST_BYTES);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          tma_load_4d_2sm(Ring + st * ST_BYTES + c * (KT * 128), &tmV, &r_full[st], ((int)rank * 2 + c) * 64, key0, b, 0);
        ++it;
      };
      for (int j = 0; j < NCH; ++j) load_k_chunk(j);  // pass 1
      load_k_chunk(0);                                // pass 2
      for (int j = 0; j < NCH; ++j) {
        const int ns = chunk_keys(j) / KT;
        if (j + 1 < NCH) load_k_chunk(j + 1);
        for (int s = 0; s < ns; ++s) load_v_tile(j * KC + s * KT);
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer: one thread of the LEADER CTA drives both tensor cores =====================
    if (lane == 0 && rank == 0) {
      const uint32_t q_addr = smem_u32(Qs), ring_addr = smem_u32(Ring), p_addr = smem_u32(Ps);
      constexpr uint32_t idescPV = make_idesc_bf16(2 * QT, HD, 0, 1);
      mbar_wait(q_full, 0);
      int it = 0;
      auto issue_S = [&](int j, int slot, int use) {
        const uint32_t idescS = make_idesc_bf16(2 * QT, chunk_keys(j), 0, 0);
        mbar_wait(&s_empty[slot], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + slot * KC;
        for (int c = 0; c < 4; ++c, ++it) {
          const int st = it % NST;
          mbar_wait(&r_full[st], (it / NST) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint64_t da = make_smem_desc_sw128(q_addr + c * (QT * 128) + kk * 32, 16, 1024);
            uint64_t db = make_smem_desc_sw128(ring_addr + st * ST_BYTES + kk * 32, 16, 1024);
            umma_bf16_2sm(d, da, db, idescS, (c | kk) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&r_empty[st]);
        }
        umma_commit_2sm(&s_full[slot]);
      };
      auto issue_PV = [&](int j, int s, uint32_t accumulate) {
        const int st = it % NST;
        mbar_wait(&p_full[s], j & 1);
        mbar_wait(&r_full[st], (it / NST) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint64_t da = make_smem_desc_sw128(p_addr + ((j * 4 + s) % NPB) * P_BYTES + kk * 32, 16, 1024);
          // half V tile: MN-major, 2 atoms of 64 dims ([64 keys x 128 B] = 8 KB apart), 16 keys per step = 2 KB
          uint64_t db = make_smem_desc_sw128(ring_addr + st * ST_BYTES + kk * (16 * 128), KT * 128, 1024);
          umma_bf16_2sm(tmem_O, da, db, idescPV, (accumulate | (uint32_t)kk) != 0 ? 1u : 0u);
        }
        umma_commit_2sm(&r_empty[st]);
        umma_commit_2sm(&pv_done[s]);
        ++it;
      };
      for (int j = 0; j < NCH; ++j) issue_S(j, j & 1, j >> 1);  // pass 1
      issue_S(0, 0, U0);                                         // pass 2
      if (NCH / 2 > 0) {
        mbar_wait(&s_empty[1], ((NCH / 2) - 1) & 1);  // O reuses the columns of S slot 1
        tc_fence_after();
      }
      uint32_t acc = 0;
      for (int j = 0; j < NCH; ++j) {
        const int ns = chunk_keys(j) / KT;
        if (j + 1 < NCH) issue_S(j + 1, 0, U0 + j + 1);
        for (int s = 0; s < ns; ++s) {
          issue_PV(j, s, acc);
          acc = 1;
        }
      }
      umma_commit_2sm(o_full);
    }
  } else if (warp < W_TMA) {
    // ===================== softmax + epilogue (identical to K1 except for the barrier arrivals) =====================
    const int wg = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const long grow = (long)q0 + r;
    const bool valid_row = grow < a.R;
    long tok = grow / a.G;
    if (tok > a.Tq - 1) tok = a.Tq - 1;
    const uint32_t* mrow = a.bits + ((long)b * a.Tq + tok) * a.W32;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    float m = -3.4e38f, l = 0.f;
    for (int j = 0; j < NCH; ++j) {
      const int slot = j & 1;
      const bool active = wg * KT < chunk_keys(j);
      const int kbase = j * KC + wg * KT;
      uint32_t mw[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
      if (active) {
        mw[0] = mrow[kbase >> 5];
        mw[1] = mrow[(kbase >> 5) + 1];
      }
      mbar_wait(&s_full[slot], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (active) {
          uint32_t sv[32];
          tmem_ld_32x32(tmem_base + lane_base + slot * KC + wg * KT + hf * 32, sv);
          tmem_ld_wait();
          const int key0 = kbase + hf * 32;
          const uint32_t w = mw[hf];
          const int nvalid = a.S_len - key0;
          if (w != 0xFFFFFFFFu) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (!((w >> c) & 1u)) sv[c] = __float_as_uint(BIG_NEG);
          }
          float tmax = -3.4e38f;
          if (nvalid >= 32) {
#pragma unroll
            for (int c = 0; c < 32; ++c) tmax = fmaxf(tmax, __uint_as_float(sv[c]));
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < nvalid) tmax = fmaxf(tmax, __uint_as_float(sv[c]));
          }
          const float m_new = fmaxf(m, tmax);
          float sum = 0.f;
          if (nvalid >= 32) {
#pragma unroll
            for (int c = 0; c < 32; ++c) sum += p_ex2((__uint_as_float(sv[c]) - m_new) * LOG2E);
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < nvalid) sum += p_ex2((__uint_as_float(sv[c]) - m_new) * LOG2E);
          }
          l = l * exp2f((m - m_new) * LOG2E) + sum;
          m = m_new;
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(&s_empty[slot], 0);
    }
    stat[(wg * 128 + r) * 2 + 0] = m;
    stat[(wg * 128 + r) * 2 + 1] = l;
    p_softmax_bar();
    {
      float mf = m;
#pragma unroll
      for (int o = 0; o < WG; ++o) mf = fmaxf(mf, stat[(o * 128 + r) * 2 + 0]);
      float lf = 0.f;
#pragma unroll
      for (int o = 0; o < WG; ++o) lf += stat[(o * 128 + r) * 2 + 1] * exp2f((stat[(o * 128 + r) * 2 + 0] - mf) * LOG2E);
      l = lf;
      m = mf;
    }
    p_softmax_bar();  // the exchange area is the P buffers
    const float inv = 1.0f / l;
    const bool leader = (warp & 3) == 0 && lane == 0;
    for (int j = 0; j < NCH; ++j) {
      const int nkeys = chunk_keys(j);
      const bool active = wg * KT < nkeys;
      const int kbase = j * KC + wg * KT;
      uint32_t mw[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
      if (active) {
        mw[0] = mrow[kbase >> 5];
        mw[1] = mrow[(kbase >> 5) + 1];
      }
      mbar_wait(&s_full[0], (U0 + j) & 1);
      tc_fence_after();
      uint32_t pk[32];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t sv[32];
        if (active) {
          tmem_ld_32x32(tmem_base + lane_base + wg * KT + hf * 32, sv);
          tmem_ld_wait();
        }
        if (hf == 1) {
          tc_fence_before();
          mbar_arrive_cluster(&s_empty[0], 0);
        }
        if (active) {
          const int key0 = kbase + hf * 32;
          const uint32_t w = mw[hf];
          const int nvalid = a.S_len - key0;
          if (w != 0xFFFFFFFFu) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (!((w >> c) & 1u)) sv[c] = __float_as_uint(BIG_NEG);
          }
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float p0 = p_ex2((__uint_as_float(sv[c]) - m) * LOG2E) * inv;
            float p1 = p_ex2((__uint_as_float(sv[c + 1]) - m) * LOG2E) * inv;
            if (c >= nvalid) p0 = 0.f;
            if (c + 1 >= nvalid) p1 = 0.f;
            pk[hf * 16 + (c >> 1)] = pack_bf16x2(p0, p1);
          }
        }
      }
      if (active) {
        const int prev = j * 4 + wg - NPB;
        if (prev >= 0) {
          mbar_wait(&pv_done[prev & 3], (prev >> 2) & 1);
          if (a.write_p) mbar_wait(&st_done[prev & 3], (prev >> 2) & 1);
        }
        uint8_t* Pbuf = Ps + ((j * 4 + wg) % NPB) * P_BYTES;
        uint8_t* prow = Pbuf + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async();
        p_wg_bar(wg);
        if (leader) {
          if (a.write_p) {
            p_tma_store_4d(&tmP, Pbuf, j * KC + wg * KT, q0, b, 0);
            p_store_commit();
          }
          mbar_arrive_cluster_release(&p_full[wg], 0);
          if (a.write_p) {
            p_store_wait_read();
            mbar_arrive(&st_done[wg]);
          }
        }
      }
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    bf16* orow = nullptr;
    if (valid_row) {
      if (grow < a.split_row) orow = a.O0 + ((long)b * a.split_row + grow) * HD;
      else orow = a.O1 + ((long)b * (a.R - a.split_row) + (grow - a.split_row)) * HD;
    }
#pragma unroll 1
    for (int c4 = 0; c4 < HD / 32 / WG; ++c4) {
      const int c = wg * (HD / 32 / WG) + c4;
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
    if (leader && a.write_p) p_store_wait_all();
  }
  tc_fence_before();
  cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer can still arrive on its barriers / read its smem
  if (warp == W_ALLOC) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

int make_tmap_bf16_4d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, int64_t s1,
                      int64_t s2, int64_t s3, uint32_t box0, uint32_t box1);  // gemm.cu

// Same contract as lapb200_fa_gemma_fwd (which forwards here when LAPB_FA_PAIR=1).
int fa_gemma_fwd_pair_launch(const void* Q, const void* Kc, const void* Vc, const uint32_t* bits, void* P, void* O0,
                             void* O1, int64_t B, int64_t R, int64_t G, int64_t Tq, int64_t S_len, int64_t Tpad,
                             int64_t W32, int64_t split_row, cudaStream_t stream) {
  using namespace pair;
  CUtensorMap tmQ, tmK, tmV, tmP;
  int rc;
  if ((rc = make_tmap_bf16_4d(&tmQ, Q, HD, R, B, 1, HD, R * HD, 0, 64, QT))) return rc;
  if ((rc = make_tmap_bf16_4d(&tmK, Kc, HD, Tpad, B, 1, HD, Tpad * HD, 0, 64, KC / 2))) return rc;  // half K slice
  if ((rc = make_tmap_bf16_4d(&tmV, Vc, HD, Tpad, B, 1, HD, Tpad * HD, 0, 64, KT))) return rc;      // one 64-dim atom
  if (P) {
    if ((rc = make_tmap_bf16_4d(&tmP, P, Tpad, R, B, 1, Tpad, R * Tpad, 0, 64, QT))) return rc;
  } else {
    tmP = tmQ;
  }
  FaPairArgs a;
  a.B = (int)B; a.R = (int)R; a.G = (int)G; a.Tq = (int)Tq; a.S_len = (int)S_len; a.Tpad = (int)Tpad;
  a.W32 = (int)W32; a.NCH = (int)((Tpad + KC - 1) / KC);
  a.bits = bits; a.O0 = (bf16*)O0; a.O1 = (bf16*)O1; a.split_row = (int)split_row; a.write_p = P ? 1 : 0;
  static bool configured = false;
  if (!configured) {
    LAPB_CUDA_OK(cudaFuncSetAttribute(fa_gemma_fwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  const unsigned tiles = (unsigned)cdiv(R, QT);
  dim3 grid((tiles + 1) & ~1u, (unsigned)B);  // an odd tile count gets one all-out-of-range tile (TMA zero-fills, rows skipped)
  fa_gemma_fwd_pair_kernel<<<grid, THREADS, SMEM, stream>>>(tmQ, tmK, tmV, tmP, a);
  LAPB_LAUNCH_OK("fa_gemma_fwd_pair");
  return 0;
}

}  // namespace lapb
