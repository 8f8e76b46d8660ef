// K1 — fused Gemma shared-attention forward for sm_100a (gemma.py:234-272), head_dim 256, one KV head.
//
//   S = Q K^T (tcgen05, accumulators in TMEM)  ->  where(mask, S, -2.3819763e38) -> softmax fp32 -> P bf16
//   O = P V   (tcgen05, P staged in shared memory as the A operand, V read MN-major straight from the cache layout)
//
// The 8 query heads of a token share the KV head, so they are stacked into the MMA M dimension: a CTA owns 128
// query rows (16 tokens x 8 heads) of one sample.  The softmax is TWO-PASS: pass 1 only accumulates the row max / sum
// (online), pass 2 recomputes S and forms p = exp(s - max) / sum rounded to bf16 — exactly the reference's rounding
// point — so the result is bit-compatible with the unfused (GEMM + softmax + GEMM) path, and P can be written out for
// the backward pass with a TMA store straight from the swizzled A-operand tile.
//
// Shape of the MMAs (v2).  A tcgen05.mma with M = 128 occupies the issue/operand path for ~128 cycles whatever its N
// (measured: the 64-key S tiles of v1 took the same time with N = 16, 32, 64 or 128), so v1's 16 N = 64 instructions per
// 64 keys left the tensor pipe 70 % idle.  Here S is produced 256 KEYS at a time: K is streamed through shared memory in
// 64-dim slices of 256 keys ([256 keys x 128 B] = 32 KB per stage, the K loop of a GEMM), 16 instructions of N = 256 per
// chunk; P V keeps its 64-key steps (4 instructions of N = 256 dims).  One ring of four 32 KB stages carries the K slices
// and the V tiles in the order the MMA warp consumes them.
//   TMEM (512 columns): pass 1 double-buffers S (2 x 256); pass 2 uses [0,256) for S and [256,512) for O.
//   softmax: four warpgroups, thread == query row == TMEM lane, warpgroup w owns keys [64w, 64w+64) of every 256-key chunk
//   — i.e. exactly one P sub-tile, which it writes (one 128-byte swizzled row per thread) when the single P buffer is free;
//   the S slices of the NEXT chunk are issued between P V (j,0) and P V (j,1) so the tensor pipe has work during the P
//   hand-offs.
// Warp roles (608 threads): warps 0-15 softmax/epilogue, warp 16 TMA producer, warp 17 MMA issuer (whole warp, elected
// lane issues), warp 18 TMEM allocator.
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace lapb {

typedef __nv_bfloat16 bf16;
#define BIG_NEG (-2.3819763e38f)

#ifndef FA_NPB
#define FA_NPB 2  // P buffers.  Measured at the training shape on one box (tools/fa_variant.py, bit-identical outputs):
#endif            // 1 buffer 273 us, 2 buffers 255 us: P V(w+1) no longer waits for P V(w) to retire before P(w+1) is written.
                  // With more than one buffer the (m, l) exchange area is aliased onto them to stay within 227 KB.
constexpr int FA_QT = 128;   // query rows per CTA
constexpr int FA_KT = 64;    // keys per P V step (= keys per softmax warpgroup and chunk)
constexpr int FA_KC = 256;   // keys per S chunk
constexpr int FA_HD = 256;   // head dim
constexpr int FA_WG = 4;     // softmax warpgroups
constexpr int FA_SOFT = 128 * FA_WG;
constexpr int FA_THREADS = FA_SOFT + 96;
constexpr int FA_NST = FA_NPB >= 4 ? 3 : 4;        // ring stages (4 P buffers leave room for 3)
constexpr int FA_Q_BYTES = FA_QT * FA_HD * 2;      // 64 KB: 4 dim-chunks of [128 rows x 128 B]
constexpr int FA_ST_BYTES = 32 * 1024;             // a K slice [256 keys x 64 dims] or a V tile [64 keys x 256 dims]
constexpr int FA_P_BYTES = FA_QT * FA_KT * 2;      // 16 KB
constexpr int FA_SMEM = FA_Q_BYTES + FA_NST * FA_ST_BYTES + FA_NPB * FA_P_BYTES + 1024 + 512 + (FA_NPB == 1 ? FA_WG * 1024 : 0);
static_assert(FA_SMEM <= 227 * 1024, "fa_gemma: shared memory");

struct FaArgs {
  int B, R, G, Tq, S_len, Tpad, W32, NCH;
  const uint32_t* bits;
  bf16* O0;
  bf16* O1;
  int split_row;  // rows [0, split_row) of a sample -> O0, the rest -> O1
  int write_p;
  int dbg;        // LAPB_FA_KNOBS builds only (tools/fa_knobs.py): bit mask of pipeline stages to skip; results are wrong
};

// Bottleneck attribution (tools/fa_knobs.py): a build with -DLAPB_FA_KNOBS can switch off one stage of the pipeline at a
// time (1 exp2 -> FMUL, 2 pass 1, 4 the P V MMAs, 8 the P shared-memory write + TMA store, 16 the pass-2 S MMAs, 32 the
// mask words).  In the product build FA_KNOB() is the constant false and all of it folds away.
#ifdef LAPB_FA_KNOBS
#define FA_KNOB(x) ((a.dbg & (x)) != 0)
#else
#define FA_KNOB(x) false
#endif
// One MUFU instruction; exp2f() without fast-math adds a range fix-up (≈3 more instructions per score) that the softmax
// does not need: inputs are <= 0, -inf gives 0, and results below 2^-126 are far below bf16's resolution of the row sum.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#ifdef FA_PRECISE_EXP2  // measurement variant (tools/fa_variant.py)
#define FA_EXP2(x) (FA_KNOB(1) ? (x) * 1e-3f : exp2f(x))
#else
#define FA_EXP2(x) (FA_KNOB(1) ? (x) * 1e-3f : ex2_approx(x))
#endif

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, %0;" ::"n"(FA_SOFT) : "memory"); }
__device__ __forceinline__ void wg_bar(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(2 + wg) : "memory"); }

// CL = CTAs per cluster (1, 2 or 4): neighbouring query tiles of one sample.  Every CTA loads 1/CL of each ring stage (a key
// range of a K slice, whole 64-dim atoms of a V tile) and multicasts it into the shared memory of all CL CTAs: the kernel
// cuts the L2 -> SM traffic (each 128-row query tile re-reads the sample's whole K twice and V once: 1.6 GB per launch at the
// training shape) by CL.  Measured: the kernel is NOT L2-bound (CL = 2 / 4 are 4 % / 14 % slower than CL = 1 because the
// CTAs of a cluster then advance in lock-step), so the launcher defaults to CL = 1; LAPB_FA_CLUSTER selects 2 or 4.
template <int CL>
__global__ void __launch_bounds__(FA_THREADS, 1)
fa_gemma_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP, const FaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;
  uint8_t* Ring = smem + FA_Q_BYTES;
  uint8_t* Ps = Ring + FA_NST * FA_ST_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ps + FA_NPB * FA_P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* r_full = bars + 1;     // [4] ring stage filled by TMA
  uint64_t* r_empty = bars + 5;    // [4] ring stage consumed by the MMAs
  uint64_t* s_full = bars + 9;     // [2] S slot written
  uint64_t* s_empty = bars + 11;   // [2] S slot read by every softmax thread
  uint64_t* p_full = bars + 13;    // [4] P sub-tile w written
  uint64_t* pv_done = bars + 17;   // [4] P V of sub-tile w retired (P buffer free again)
  uint64_t* st_done = bars + 21;   // [4] TMA store of sub-tile w has read the P buffer
  uint64_t* o_full = bars + 25;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  // [FA_WG][128][2] (m, l) exchange between the warpgroups (with two P buffers it borrows them: P is not written before)
  float* stat = FA_NPB == 1 ? reinterpret_cast<float*>(bars + 64) : reinterpret_cast<float*>(Ps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * FA_QT;
  const int NCH = a.NCH;
  constexpr int W_TMA = 4 * FA_WG, W_MMA = W_TMA + 1, W_ALLOC = W_TMA + 2;
  const int U0 = FA_KNOB(2) ? 0 : (NCH + 1) / 2;  // uses of S slot 0 in pass 1

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == W_MMA && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < FA_NST; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], CL);  // released by the MMA warp of every CTA that received the multicast stage
    }
    for (int i = 0; i < FA_WG; ++i) {
      mbar_init(&p_full[i], 1);
      mbar_init(&pv_done[i], 1);
      mbar_init(&st_done[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], FA_SOFT);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == W_ALLOC) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peers' barriers must exist before anything is multicast into them
  tc_fence_after();
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;  // pass 2: S in columns [0,256), O in [256,512)
  auto chunk_keys = [&](int j) { return min(FA_KC, a.Tpad - j * FA_KC); };  // multiple of 64

  if (warp == W_TMA) {
    // ===================== TMA producer: one ring, loads in the order the MMA warp consumes them =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, FA_Q_BYTES);
#pragma unroll
      for (int c = 0; c < 4; ++c) tma_load_4d(Qs + c * (FA_QT * 128), &tmQ, q_full, c * 64, q0, b, 0);
      int it = 0;
      auto load_k_chunk = [&](int j) {  // four 64-dim slices of 256 keys
        for (int c = 0; c < 4; ++c, ++it) {
          const int st = it % FA_NST;
          mbar_wait(&r_empty[st], ((it / FA_NST) & 1) ^ 1);
          mbar_expect_tx(&r_full[st], FA_ST_BYTES);
          if (CL > 1)  // this CTA's key range of the slice, to every CTA of the cluster
            tma_load_4d_mc(Ring + st * FA_ST_BYTES + crank * (FA_KC / CL) * 128, &tmK, &r_full[st], c * 64,
                           j * FA_KC + crank * (FA_KC / CL), b, 0, CMASK);
          else
            tma_load_4d(Ring + st * FA_ST_BYTES, &tmK, &r_full[st], c * 64, j * FA_KC, b, 0);
        }
      };
      auto load_v_tile = [&](int key0) {  // [64 keys x 256 dims]: 4 atoms of 64 dims ([64 keys x 128 B] each)
        const int st = it % FA_NST;
        mbar_wait(&r_empty[st], ((it / FA_NST) & 1) ^ 1);
        mbar_expect_tx(&r_full[st], FA_ST_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (CL > 1) {  // atoms c = crank, crank + CL, ... to every CTA of the cluster
            if ((c % CL) == (int)crank)
              tma_load_4d_mc(Ring + st * FA_ST_BYTES + c * (FA_KT * 128), &tmV, &r_full[st], c * 64, key0, b, 0, CMASK);
          } else {
            tma_load_4d(Ring + st * FA_ST_BYTES + c * (FA_KT * 128), &tmV, &r_full[st], c * 64, key0, b, 0);
          }
        }
        ++it;
      };
      if (!FA_KNOB(2))
        for (int j = 0; j < NCH; ++j) load_k_chunk(j);  // pass 1
      load_k_chunk(0);                                // pass 2
      for (int j = 0; j < NCH; ++j) {
        const int ns = chunk_keys(j) / FA_KT;
        if (j + 1 < NCH) load_k_chunk(j + 1);
        for (int s = 0; s < ns; ++s) load_v_tile(j * FA_KC + s * FA_KT);
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (whole warp; an elected lane issues) =====================
    const uint32_t q_addr = smem_u32(Qs), ring_addr = smem_u32(Ring), p_addr = smem_u32(Ps);
    constexpr uint32_t idescPV = make_idesc_bf16(FA_QT, FA_HD, 0, 1);
    mbar_wait(q_full, 0);
    int it = 0;
    auto issue_S = [&](int j, int slot, int use, bool skip = false) {
      const uint32_t idescS = make_idesc_bf16(FA_QT, chunk_keys(j), 0, 0);
      mbar_wait(&s_empty[slot], (use & 1) ^ 1);
      const uint32_t d = tmem_base + slot * FA_KC;
      for (int c = 0; c < 4; ++c, ++it) {
        const int st = it % FA_NST;
        mbar_wait(&r_full[st], (it / FA_NST) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint64_t da = make_smem_desc_sw128(q_addr + c * (FA_QT * 128) + kk * 32, 16, 1024);
          uint64_t db = make_smem_desc_sw128(ring_addr + st * FA_ST_BYTES + kk * 32, 16, 1024);
          if (!skip) umma_bf16_elect(d, da, db, idescS, (c | kk) != 0 ? 1u : 0u);
        }
        if (CL > 1) umma_commit_mc_elect(&r_empty[st], CMASK);
        else umma_commit_elect(&r_empty[st]);
      }
      umma_commit_elect(&s_full[slot]);
    };
    auto issue_PV = [&](int j, int s, uint32_t accumulate) {
      const int st = it % FA_NST;
      mbar_wait(&p_full[s], j & 1);
      mbar_wait(&r_full[st], (it / FA_NST) & 1);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t da = make_smem_desc_sw128(p_addr + ((j * 4 + s) % FA_NPB) * FA_P_BYTES + kk * 32, 16, 1024);
        // V tile: MN-major, 4 atoms of 64 dims ([64 keys x 128 B] = 8 KB apart), 16 keys per step = 2 KB
        uint64_t db = make_smem_desc_sw128(ring_addr + st * FA_ST_BYTES + kk * (16 * 128), FA_KT * 128, 1024);
        if (!FA_KNOB(4)) umma_bf16_elect(tmem_O, da, db, idescPV, (accumulate | (uint32_t)kk) != 0 ? 1u : 0u);
      }
      if (CL > 1) umma_commit_mc_elect(&r_empty[st], CMASK);
      else umma_commit_elect(&r_empty[st]);
      umma_commit_elect(&pv_done[s]);
      ++it;
    };
    if (!FA_KNOB(2))
      for (int j = 0; j < NCH; ++j) issue_S(j, j & 1, j >> 1);  // pass 1: logits only, S double-buffered
    issue_S(0, 0, U0, FA_KNOB(16));                             // pass 2
    if (NCH / 2 > 0 && !FA_KNOB(2)) mbar_wait(&s_empty[1], ((NCH / 2) - 1) & 1);  // O reuses the columns of S slot 1
    uint32_t acc = 0;
    for (int j = 0; j < NCH; ++j) {
      const int ns = chunk_keys(j) / FA_KT;
      // S(j+1) goes first: it runs on the tensor pipe while the softmax warps spend their ~2 K MUFU cycles on chunk j
      // (its slot is free as soon as every softmax thread has pulled S(j) into registers)
      if (j + 1 < NCH) issue_S(j + 1, 0, U0 + j + 1, FA_KNOB(16));
      for (int s = 0; s < ns; ++s) {
        issue_PV(j, s, acc);
        acc = 1;
      }
    }
    umma_commit_elect(o_full);
  } else if (warp < W_TMA) {
    // ===================== softmax + epilogue: thread == (query row, 64-key slice of every chunk) =====================
    const int wg = warp >> 2;
    const int r = (warp & 3) * 32 + lane;     // query row of the tile == TMEM lane
    const long grow = (long)q0 + r;
    const bool valid_row = grow < a.R;
    long tok = grow / a.G;
    if (tok > a.Tq - 1) tok = a.Tq - 1;
    const uint32_t* mrow = a.bits + ((long)b * a.Tq + tok) * a.W32;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    float m = -3.4e38f, l = 0.f;
    // ---- pass 1: running max / sum over this warpgroup's keys ----
    for (int j = 0; j < (FA_KNOB(2) ? 0 : NCH); ++j) {
      const int slot = j & 1;
      const bool active = wg * FA_KT < chunk_keys(j);
      // the two mask words of this thread's 64 keys are fetched BEFORE the wait, so their latency hides behind the MMAs
      const int kbase = j * FA_KC + wg * FA_KT;
      uint32_t mw[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
      if (active && !FA_KNOB(32)) {
        mw[0] = mrow[kbase >> 5];
        mw[1] = mrow[(kbase >> 5) + 1];
      }
      mbar_wait(&s_full[slot], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (active) {
          uint32_t sv[32];
          tmem_ld_32x32(tmem_base + lane_base + slot * FA_KC + wg * FA_KT + hf * 32, sv);
          tmem_ld_wait();
          const int key0 = j * FA_KC + wg * FA_KT + hf * 32;
          const uint32_t w = mw[hf];
          const int nvalid = a.S_len - key0;  // columns [0, nvalid) are real keys
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
          // (x - m) is formed BEFORE scaling by log2(e): for a fully masked row m = -2.38e38 and m*log2(e) would overflow
          if (nvalid >= 32) {
#pragma unroll
            for (int c = 0; c < 32; ++c) sum += FA_EXP2((__uint_as_float(sv[c]) - m_new) * LOG2E);
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < nvalid) sum += FA_EXP2((__uint_as_float(sv[c]) - m_new) * LOG2E);
          }
          l = l * exp2f((m - m_new) * LOG2E) + sum;
          m = m_new;
        }
      }
      tc_fence_before();
      mbar_arrive_relaxed(&s_empty[slot]);
    }
    // combine the key slices of the warpgroups
    if (FA_KNOB(2)) { m = 20.f; l = 1000.f; }
    stat[(wg * 128 + r) * 2 + 0] = m;
    stat[(wg * 128 + r) * 2 + 1] = l;
    softmax_bar();
    {
      float mf = m;
#pragma unroll
      for (int o = 0; o < FA_WG; ++o) mf = fmaxf(mf, stat[(o * 128 + r) * 2 + 0]);
      float lf = 0.f;
#pragma unroll
      for (int o = 0; o < FA_WG; ++o) lf += stat[(o * 128 + r) * 2 + 1] * exp2f((stat[(o * 128 + r) * 2 + 0] - mf) * LOG2E);
      l = lf;
      m = mf;
    }
    if (FA_NPB > 1) softmax_bar();  // every thread has read the exchange area before P sub-tiles overwrite it
    // ---- pass 2: p = exp(s - max) / sum -> bf16 -> this warpgroup's P sub-tile [+ TMA store for the backward] ----
    const float inv = 1.0f / l;
    const bool leader = (warp & 3) == 0 && lane == 0;
    for (int j = 0; j < NCH; ++j) {
      const int nkeys = chunk_keys(j);
      const bool active = wg * FA_KT < nkeys;
      const int kbase = j * FA_KC + wg * FA_KT;
      uint32_t mw[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
      if (active && !FA_KNOB(32)) {
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
          tmem_ld_32x32(tmem_base + lane_base + wg * FA_KT + hf * 32, sv);
          tmem_ld_wait();
        }
        if (hf == 1) {  // S(j) is out of TMEM: the next chunk's S may overwrite the slot while the exps below run
          tc_fence_before();
          mbar_arrive_relaxed(&s_empty[0]);
        }
        if (active) {
          const int key0 = j * FA_KC + wg * FA_KT + hf * 32;
          const uint32_t w = mw[hf];
          const int nvalid = a.S_len - key0;
          if (w != 0xFFFFFFFFu) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (!((w >> c) & 1u)) sv[c] = __float_as_uint(BIG_NEG);
          }
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float p0 = FA_EXP2((__uint_as_float(sv[c]) - m) * LOG2E) * inv;
            float p1 = FA_EXP2((__uint_as_float(sv[c + 1]) - m) * LOG2E) * inv;
            if (c >= nvalid) p0 = 0.f;
            if (c + 1 >= nvalid) p1 = 0.f;
            pk[hf * 16 + (c >> 1)] = pack_bf16x2(p0, p1);
          }
        }
      }
      if (active) {
        // P buffer (sub-tile n) % FA_NPB is free once the P V of sub-tile n - FA_NPB has retired (and its TMA store has read
        // it); every chunk before the last has all four sub-tiles, so sub-tile n = 4 j + wg
        const int prev = j * 4 + wg - FA_NPB;
        if (prev >= 0) {
          mbar_wait(&pv_done[prev & 3], (prev >> 2) & 1);
          if (a.write_p) mbar_wait(&st_done[prev & 3], (prev >> 2) & 1);
        }
        uint8_t* Pbuf = Ps + ((j * 4 + wg) % FA_NPB) * FA_P_BYTES;
        // K-major, 128B-swizzled A tile: row r is 128 B (64 keys); 16-byte chunk c sits at chunk position c ^ (r & 7)
        uint8_t* prow = Pbuf + r * 128;
        if (!FA_KNOB(8)) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        fence_proxy_async();
        wg_bar(wg);
        if (leader) {
          if (a.write_p && !FA_KNOB(8)) {
            tma_store_4d(&tmP, Pbuf, j * FA_KC + wg * FA_KT, q0, b, 0);
            tma_store_commit();
          }
          mbar_arrive(&p_full[wg]);
          if (a.write_p) {
            tma_store_wait_read();
            mbar_arrive(&st_done[wg]);
          }
        }
      }
    }
    // ---- epilogue: O (fp32, TMEM) -> bf16 rows; warpgroup g stores dims [64g, 64g + 64) ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    bf16* orow = nullptr;
    if (valid_row) {
      if (grow < a.split_row) orow = a.O0 + ((long)b * a.split_row + grow) * FA_HD;
      else orow = a.O1 + ((long)b * (a.R - a.split_row) + (grow - a.split_row)) * FA_HD;
    }
#pragma unroll 1
    for (int c4 = 0; c4 < FA_HD / 32 / FA_WG; ++c4) {
      const int c = wg * (FA_HD / 32 / FA_WG) + c4;
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
    if (leader && a.write_p) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer can still arrive on its barriers
  if (warp == W_ALLOC) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int make_tmap_bf16_4d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, int64_t s1,
                      int64_t s2, int64_t s3, uint32_t box0, uint32_t box1);  // gemm.cu

int fa_gemma_fwd_pair_launch(const void* Q, const void* Kc, const void* Vc, const uint32_t* bits, void* P, void* O0,
                             void* O1, int64_t B, int64_t R, int64_t G, int64_t Tq, int64_t S_len, int64_t Tpad,
                             int64_t W32, int64_t split_row, cudaStream_t stream);  // fa_gemma_pair.cu (experimental)

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
  static int pair_env = -1;
  if (pair_env < 0) {
    const char* e = getenv("LAPB_FA_PAIR");
    pair_env = e ? atoi(e) : 0;  // experimental cta_group::2 kernel (fa_gemma_pair.cu), off by default
  }
  if (pair_env) return fa_gemma_fwd_pair_launch(Q, Kc, Vc, bits, P, O0, O1, B, R, G, Tq, S_len, Tpad, W32, split_row, stream);
  CUtensorMap tmQ, tmK, tmV, tmP;
  int rc;
  if ((rc = make_tmap_bf16_4d(&tmQ, Q, FA_HD, R, B, 1, FA_HD, R * FA_HD, 0, 64, FA_QT))) return rc;
  dim3 grid(cdiv(R, FA_QT), (unsigned)B);
  static int cl_env = -1;
  if (cl_env < 0) {
    const char* e = getenv("LAPB_FA_CLUSTER");
    cl_env = e ? atoi(e) : 1;  // measured at the training shape: CL = 1 282 us, CL = 2 292 us, CL = 4 321 us
  }
  const int CLN = (cl_env >= 4 && grid.x % 4 == 0) ? 4 : (cl_env >= 2 && grid.x % 2 == 0) ? 2 : 1;
  if ((rc = make_tmap_bf16_4d(&tmK, Kc, FA_HD, Tpad, B, 1, FA_HD, Tpad * FA_HD, 0, 64, FA_KC / CLN))) return rc;  // K slice
  if ((rc = make_tmap_bf16_4d(&tmV, Vc, FA_HD, Tpad, B, 1, FA_HD, Tpad * FA_HD, 0, 64, FA_KT))) return rc;
  if (P) {
    if ((rc = make_tmap_bf16_4d(&tmP, P, Tpad, R, B, 1, Tpad, R * Tpad, 0, 64, FA_QT))) return rc;
  } else {
    tmP = tmQ;
  }
  FaArgs a;
  a.B = (int)B; a.R = (int)R; a.G = (int)G; a.Tq = (int)Tq; a.S_len = (int)S_len; a.Tpad = (int)Tpad;
  a.W32 = (int)W32; a.NCH = (int)((Tpad + FA_KC - 1) / FA_KC);
  a.bits = bits; a.O0 = (bf16*)O0; a.O1 = (bf16*)O1; a.split_row = (int)split_row; a.write_p = P ? 1 : 0;
  a.dbg = 0;
#ifdef LAPB_FA_KNOBS
  if (const char* e = getenv("LAPB_FA_KNOBS")) a.dbg = atoi(e);
#endif
  static bool configured = false;
  if (!configured) {
    LAPB_CUDA_OK(cudaFuncSetAttribute(fa_gemma_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    LAPB_CUDA_OK(cudaFuncSetAttribute(fa_gemma_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    LAPB_CUDA_OK(cudaFuncSetAttribute(fa_gemma_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    configured = true;
  }
  if (CLN > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(FA_THREADS);
    cfg.dynamicSmemBytes = FA_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLN;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (CLN == 4) LAPB_CUDA_OK(cudaLaunchKernelEx(&cfg, fa_gemma_fwd_kernel<4>, tmQ, tmK, tmV, tmP, a));
    else LAPB_CUDA_OK(cudaLaunchKernelEx(&cfg, fa_gemma_fwd_kernel<2>, tmQ, tmK, tmV, tmP, a));
  } else {
    fa_gemma_fwd_kernel<1><<<grid, FA_THREADS, FA_SMEM, stream>>>(tmQ, tmK, tmV, tmP, a);
  }
  LAPB_LAUNCH_OK("fa_gemma_fwd");
  return 0;
}
