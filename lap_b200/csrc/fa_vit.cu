// K2 — fused SigLIP (ViT) attention forward for sm_100a: flax MultiHeadDotProductAttention of OP/models/siglip.py:88-93
// between the QKV projection and the output projection.
//
//   S = Q K^T   (Q already divided by sqrt(head_dim) by the QKV GEMM's epilogue; tcgen05, accumulator in TMEM)
//   P = softmax(S) with the reference's rounding points (bf16 logits; mode 0: the softmax itself in bf16 — flax 0.10.2
//       evaluates it in the compute dtype — mode 1: fp32 softmax, one rounding of P)
//   O = P V     (tcgen05; P staged in shared memory as the A operand, V read MN-major straight from the qkv rows)
//
// One CTA = 128 query rows of one (image, head).  head_dim = 72 is neither a swizzle atom nor a multiple of 16, so the
// tensor maps describe the head as a 72-wide innermost dimension and the boxes are 64 wide: the second box of every operand
// covers dims [64, 128), of which [72, 128) are out of bounds and arrive as ZEROS from the TMA unit — the padding lives in
// shared memory only, HBM keeps [tokens, 3, heads, 72].  S needs 5 K-steps of 16 dims (64 + 16), P V uses N = 128 columns
// of which 72 are stored.  Keys are processed 256 at a time (the whole row for 224 px = 256 patches; 3 chunks for 384 px =
// 729 patches).  P is written OVER the K chunk (K is dead once S is in TMEM), so Q 32 KB + K/P 64 KB + V 64 KB fit one SM and
// Q, K and V are all requested at kernel start.
//
// Softmax: four warpgroups, thread == query row == TMEM lane, warpgroup w owns keys [64w, 64w + 64) of a chunk == one P
// sub-tile [128 x 64] written as 128-byte swizzled rows (A operand of P V) and TMA-stored to HBM for the backward pass.
// With one chunk the row statistics come from the single S in TMEM (exact bf16 emulation: max, then e = bf16(exp(bf16(s -
// max))), sum, bf16(sum), p = bf16(e / sum)); with several chunks pass 1 accumulates max / sum online and pass 2 recomputes
// S — the same two-pass scheme as K1 (fa_gemma.cu).
// Warp roles (608 threads): warps 0-15 softmax / epilogue, warp 16 TMA producer, warp 17 MMA issuer, warp 18 TMEM allocator.
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace lapb {

typedef __nv_bfloat16 bf16;
#define FV_NEG_INF (__int_as_float(0xff800000))

constexpr int FV_QT = 128;                    // query rows per CTA
constexpr int FV_KC = 256;                    // keys per chunk
constexpr int FV_KT = 64;                     // keys per P sub-tile (= per softmax warpgroup)
constexpr int FV_HP = 128;                    // padded head dim in shared memory (2 atoms of 64)
constexpr int FV_WG = 4;
constexpr int FV_SOFT = 128 * FV_WG;
constexpr int FV_THREADS = FV_SOFT + 96;
constexpr int FV_Q_BYTES = FV_QT * FV_HP * 2;   // 32 KB: 2 atoms [128 rows x 128 B]
constexpr int FV_KV_BYTES = FV_KC * FV_HP * 2;  // 64 KB: 2 atoms [256 keys x 128 B]
constexpr int FV_P_BYTES = FV_QT * FV_KC * 2;   // 64 KB: 4 sub-tiles [128 rows x 128 B] — aliases the K chunk
static_assert(FV_P_BYTES <= FV_KV_BYTES, "P must fit over the K chunk");
constexpr int FV_SMEM = FV_Q_BYTES + 2 * FV_KV_BYTES + 1024 + 512 + FV_WG * 128 * 8;
static_assert(FV_SMEM <= 227 * 1024, "fa_vit: shared memory");

struct FvArgs {
  int Np, NCH, hd, mode, write_p;
  long ldo;   // row stride of O (elements): rows are (image, token), head h occupies columns [h*hd, h*hd + hd)
  bf16* O;
};

__device__ __forceinline__ void fv_tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void fv_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fv_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fv_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fv_softmax_bar() { asm volatile("bar.sync 1, %0;" ::"n"(FV_SOFT) : "memory"); }
__device__ __forceinline__ void fv_wg_bar(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(2 + wg) : "memory"); }

__global__ void __launch_bounds__(FV_THREADS, 1)
fa_vit_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP, const FvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;
  uint8_t* Ks = Qs + FV_Q_BYTES;
  uint8_t* Ps = Ks;               // P overwrites the K chunk once S is in TMEM
  uint8_t* Vs = Ks + FV_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + FV_KV_BYTES);
  uint64_t* q_full = bars;        // Q tile landed
  uint64_t* k_full = bars + 1;    // a K chunk landed
  uint64_t* k_free = bars + 2;    // the last MMAs reading the K / P buffer have retired (S in pass 1, P V in pass 2)
  uint64_t* v_full = bars + 15;   // a V chunk landed
  uint64_t* v_free = bars + 16;   // the P V MMAs reading the V chunk have retired
  uint64_t* s_full = bars + 3;    // S of a chunk is in TMEM
  uint64_t* s_free = bars + 4;    // every softmax thread has pulled its S values
  uint64_t* p_full = bars + 5;    // [4] P sub-tile w written
  uint64_t* pv_done = bars + 9;   // P V of a chunk retired (P buffer free)
  uint64_t* st_done = bars + 10;  // [4] the TMA store of sub-tile w has read the P buffer
  uint64_t* o_full = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  float* stat = reinterpret_cast<float*>(bars + 64);  // [FV_WG][128][2] exchange between the warpgroups

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * FV_QT, h = blockIdx.y, img = blockIdx.z;
  const int NCH = a.NCH;
  const bool two_pass = NCH > 1;
  constexpr int W_TMA = 4 * FV_WG, W_MMA = W_TMA + 1, W_ALLOC = W_TMA + 2;
  auto chunk_keys = [&](int j) { return min(FV_KC, ((a.Np - j * FV_KC) + 15) & ~15); };  // multiple of 16

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == W_MMA && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_free, 1);
    mbar_init(v_full, 1);
    mbar_init(v_free, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, FV_SOFT);
    for (int i = 0; i < FV_WG; ++i) {
      mbar_init(&p_full[i], 1);
      mbar_init(&st_done[i], 1);
    }
    mbar_init(pv_done, 1);
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
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;  // S in columns [0, 256), O in [256, 384)

  if (warp == W_TMA) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, FV_Q_BYTES);
      tma_load_4d(Qs, &tmQ, q_full, 0, q0, h, img);
      tma_load_4d(Qs + FV_QT * 128, &tmQ, q_full, 64, q0, h, img);  // dims [64, 128): >= 72 arrive as zeros
      int ku = 0;  // K loads so far
      auto load_k = [&](int j) {
        mbar_wait(k_free, (ku & 1) ^ 1);
        mbar_expect_tx(k_full, FV_KV_BYTES);
        tma_load_4d(Ks, &tmK, k_full, 0, j * FV_KC, h, img);
        tma_load_4d(Ks + FV_KC * 128, &tmK, k_full, 64, j * FV_KC, h, img);
        ++ku;
      };
      if (two_pass)
        for (int j = 0; j < NCH; ++j) load_k(j);
      for (int j = 0; j < NCH; ++j) {
        if (j > 0 && a.write_p)  // the TMA stores of P(j-1) must have read the buffer K(j) is about to overwrite
          for (int w = 0; w < FV_WG; ++w) mbar_wait(&st_done[w], (j - 1) & 1);
        load_k(j);
        mbar_wait(v_free, (j & 1) ^ 1);
        mbar_expect_tx(v_full, FV_KV_BYTES);
        tma_load_4d(Vs, &tmV, v_full, 0, j * FV_KC, h, img);
        tma_load_4d(Vs + FV_KC * 128, &tmV, v_full, 64, j * FV_KC, h, img);
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (whole warp; an elected lane issues) =====================
    const uint32_t q_addr = smem_u32(Qs), k_addr = smem_u32(Ks), v_addr = smem_u32(Vs), p_addr = smem_u32(Ps);
    constexpr uint32_t idescPV = make_idesc_bf16(FV_QT, FV_HP, 0, 1);
    mbar_wait(q_full, 0);
    int ku = 0, s_use = 0;
    auto issue_S = [&](int j, bool release_k) {
      const uint32_t idescS = make_idesc_bf16(FV_QT, chunk_keys(j), 0, 0);
      mbar_wait(k_full, ku & 1);
      mbar_wait(s_free, (s_use & 1) ^ 1);
      tc_fence_after();
      // dims [0, 64): four 16-dim steps inside atom 0; dims [64, 80): one step inside atom 1 (72..79 are zero)
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) {
        const uint32_t qo = (kk < 4) ? kk * 32 : FV_QT * 128;
        const uint32_t ko = (kk < 4) ? kk * 32 : FV_KC * 128;
        uint64_t da = make_smem_desc_sw128(q_addr + qo, 16, 1024);
        uint64_t db = make_smem_desc_sw128(k_addr + ko, 16, 1024);
        umma_bf16_elect(tmem_base, da, db, idescS, kk != 0 ? 1u : 0u);
      }
      if (release_k) umma_commit_elect(k_free);  // pass 1: K is free again; pass 2: P goes there, freed after P V
      umma_commit_elect(s_full);
      ++ku;
      ++s_use;
    };
    if (two_pass)
      for (int j = 0; j < NCH; ++j) issue_S(j, true);
    for (int j = 0; j < NCH; ++j) {
      issue_S(j, false);
      mbar_wait(v_full, j & 1);  // V of chunk j
      const int ns = (chunk_keys(j) + FV_KT - 1) / FV_KT;
      for (int s = 0; s < ns; ++s) {
        mbar_wait(&p_full[s], j & 1);
        tc_fence_after();
        const int ksteps = min(FV_KT, chunk_keys(j) - s * FV_KT) / 16;
        for (int kk = 0; kk < ksteps; ++kk) {
          uint64_t da = make_smem_desc_sw128(p_addr + s * (FV_QT * 128) + kk * 32, 16, 1024);
          // V chunk: MN-major, 2 atoms of 64 dims ([256 keys x 128 B] = 32 KB apart), 16 keys per step = 2 KB
          uint64_t db = make_smem_desc_sw128(v_addr + (s * FV_KT + kk * 16) * 128, FV_KC * 128, 1024);
          umma_bf16_elect(tmem_O, da, db, idescPV, (j | s | kk) != 0 ? 1u : 0u);
        }
      }
      umma_commit_elect(k_free);
      umma_commit_elect(v_free);
      umma_commit_elect(pv_done);
    }
    umma_commit_elect(o_full);
  } else if (warp < W_TMA) {
    // ===================== softmax + epilogue =====================
    const int wg = warp >> 2;
    const int r = (warp & 3) * 32 + lane;  // query row of the tile == TMEM lane
    const int qrow = q0 + r;
    const bool valid_row = qrow < a.Np;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const bool bf16_softmax = a.mode == 0;
    const bool leader = (warp & 3) == 0 && lane == 0;
    float m = -3.4e38f, l = 0.f;
    int s_use = 0;
    // this thread's 64 logits of a chunk, rounded to bf16 (the reference's logits are a bf16 array); keys >= Np -> -inf
    auto load_scores = [&](int j, float (&sv)[64]) {
      const int k0 = j * FV_KC + wg * FV_KT;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t t[32];
        tmem_ld_32x32(tmem_base + lane_base + wg * FV_KT + hf * 32, t);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c)
          sv[hf * 32 + c] = (k0 + hf * 32 + c < a.Np) ? bf16r(__uint_as_float(t[c])) : FV_NEG_INF;
      }
    };
    auto release_S = [&]() {
      tc_fence_before();
      mbar_arrive_relaxed(s_free);
      ++s_use;
    };
    auto expv = [&](float x) {  // exp of (s - max) with the reference's roundings
      return bf16_softmax ? bf16r(__expf(bf16r(x))) : __expf(x);
    };
    auto exchange = [&](float mine, bool is_max) -> float {  // combine one value per warpgroup across the four of them
      stat[(wg * 128 + r) * 2] = mine;
      fv_softmax_bar();
      float v = stat[r * 2];
#pragma unroll
      for (int o = 1; o < FV_WG; ++o) v = is_max ? fmaxf(v, stat[(o * 128 + r) * 2]) : v + stat[(o * 128 + r) * 2];
      fv_softmax_bar();
      return v;
    };
    float sv[64];
    if (two_pass) {
      // ---- pass 1: running max / sum over this warpgroup's keys of every chunk ----
      for (int j = 0; j < NCH; ++j) {
        const bool active = wg * FV_KT < chunk_keys(j);
        mbar_wait(s_full, s_use & 1);
        tc_fence_after();
        if (active) load_scores(j, sv);
        release_S();
        if (active) {
          float tmax = m;
#pragma unroll
          for (int c = 0; c < 64; ++c) tmax = fmaxf(tmax, sv[c]);
          float sum = 0.f;
#pragma unroll
          for (int c = 0; c < 64; ++c) sum += expv(sv[c] - tmax);
          l = l * __expf(m - tmax) + sum;
          m = tmax;
        }
      }
      const float mf = exchange(m, true);
      l = exchange(l * __expf(m - mf), false);
      m = mf;
    }
    for (int j = 0; j < NCH; ++j) {
      const int nkeys = chunk_keys(j);
      const bool active = wg * FV_KT < nkeys;
      mbar_wait(s_full, s_use & 1);
      tc_fence_after();
      if (active) {
        load_scores(j, sv);
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) sv[c] = FV_NEG_INF;
      }
      release_S();
      if (!two_pass) {
        // the whole row is in this one chunk: exact statistics, in the reference's order (max, exp, sum)
        float tmax = -3.4e38f;
#pragma unroll
        for (int c = 0; c < 64; ++c) tmax = fmaxf(tmax, sv[c]);
        m = exchange(tmax, true);
#pragma unroll
        for (int c = 0; c < 64; ++c) sv[c] = expv(sv[c] - m);
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) sum += sv[c];
        l = exchange(sum, false);
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) sv[c] = expv(sv[c] - m);
      }
      const float inv = 1.0f / (bf16_softmax ? bf16r(l) : l);  // see the pair kernel
      if (active) {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 64; c += 2) pk[c >> 1] = pack_bf16x2(sv[c] * inv, sv[c + 1] * inv);
        // (P aliases K(j): the producer loaded K(j) only after P V(j-1) had retired and the stores of P(j-1) had read the
        //  buffer, and s_full(j) says the S MMAs are done reading K(j) — the buffer is ours)
        // K-major, 128B-swizzled A sub-tile: row r is 128 B (64 keys); 16-byte chunk c sits at chunk position c ^ (r & 7)
        uint8_t* prow = Ps + wg * (FV_QT * 128) + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async();
        fv_wg_bar(wg);
        if (leader) {
          if (a.write_p) {
            fv_tma_store_4d(&tmP, Ps + wg * (FV_QT * 128), j * FV_KC + wg * FV_KT, q0, h, img);
            fv_store_commit();
          }
          mbar_arrive(&p_full[wg]);
          if (a.write_p) {
            fv_store_wait_read();
            mbar_arrive(&st_done[wg]);
          }
        }
      }
    }
    // ---- epilogue: O (fp32, TMEM) -> bf16 rows of hd valid columns ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    if (wg * 32 < a.hd) {  // warpgroup g stores columns [32 g, 32 g + 32) of its row
      bf16* orow = a.O + ((long)img * a.Np + qrow) * a.ldo + (long)h * a.hd;
      {
        const int c0 = wg * 32;
        uint32_t o[32];
        tmem_ld_32x32(tmem_O + lane_base + c0, o);
        tmem_ld_wait();
        if (valid_row) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if (c0 + v * 8 < a.hd) {  // hd % 8 == 0
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(o[8 * v + 0]), __uint_as_float(o[8 * v + 1]));
              u.y = pack_bf16x2(__uint_as_float(o[8 * v + 2]), __uint_as_float(o[8 * v + 3]));
              u.z = pack_bf16x2(__uint_as_float(o[8 * v + 4]), __uint_as_float(o[8 * v + 5]));
              u.w = pack_bf16x2(__uint_as_float(o[8 * v + 6]), __uint_as_float(o[8 * v + 7]));
              *reinterpret_cast<uint4*>(orow + c0 + v * 8) = u;
            }
          }
        }
      }
    }
    if (leader && a.write_p) fv_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_ALLOC) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Single-chunk variant (Np <= 256: the 224 px tower): ONE CTA = one (image, head) = BOTH 128-row query tiles, so K and V are
// fetched once per head and the two tiles are pipelined through the roles — while the softmax warps work on tile 1 the
// tensor core runs P V of tile 0, and the epilogue of tile 0 overlaps P V of tile 1.  (The one-tile-per-CTA kernel above is
// latency-bound: ncu shows 30 % issue utilisation with the CTA's phases strictly serial, profiles/r02_k2.md.)
//   shared memory: Q0 | Q1 (2 x 32 KB), K (64 KB), V (64 KB).  P(0) overwrites K, P(1) overwrites Q0|Q1: both are dead once
//   the two S MMAs have retired.  TMEM: S0 [0, 256), S1 [256, 512); O(t) reuses the first 128 columns of S(t).
// ------------------------------------------------------------------------------------------------------------------
constexpr int FVP_SMEM = 2 * FV_Q_BYTES + 2 * FV_KV_BYTES + 1024 + 512 + 4 * FV_WG * 128 * 4;  // 4 exchange slots
static_assert(FVP_SMEM <= 227 * 1024, "fa_vit pair: shared memory");

__global__ void __launch_bounds__(FV_THREADS, 1)
fa_vit_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP, const FvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;                       // [2 tiles][2 atoms][128 rows x 128 B]
  uint8_t* Ks = Qs + 2 * FV_Q_BYTES;
  uint8_t* Vs = Ks + FV_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + FV_KV_BYTES);
  uint64_t* qk_full = bars;        // Q0, Q1 and K landed
  uint64_t* v_full = bars + 1;
  uint64_t* s_full = bars + 2;     // [2] S(t) in TMEM (implies every S MMA issued before it has retired)
  uint64_t* s_read = bars + 4;     // [2] every softmax thread has pulled its S(t) values: O(t) may overwrite the columns
  uint64_t* p_full = bars + 6;     // [2][4] P sub-tile w of tile t written
  uint64_t* o_full = bars + 14;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* stat = reinterpret_cast<float*>(bars + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, img = blockIdx.y;
  const int NT = a.Np > FV_QT ? 2 : 1;                      // query tiles of this head
  const int nkeys = (a.Np + 15) & ~15;                      // MMA N / K extent (multiple of 16, <= 256)
  constexpr int W_TMA = 4 * FV_WG, W_MMA = W_TMA + 1, W_ALLOC = W_TMA + 2;
  uint8_t* const Pbuf[2] = {Ks, Qs};

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == W_MMA && lane == 0) {
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_read[t], FV_SOFT);
      mbar_init(&o_full[t], 1);
      for (int w = 0; w < FV_WG; ++w) mbar_init(&p_full[t * 4 + w], 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == W_ALLOC) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    if (lane == 0) {
      mbar_expect_tx(qk_full, NT * FV_Q_BYTES + FV_KV_BYTES);
      for (int t = 0; t < NT; ++t) {
        tma_load_4d(Qs + t * FV_Q_BYTES, &tmQ, qk_full, 0, t * FV_QT, h, img);
        tma_load_4d(Qs + t * FV_Q_BYTES + FV_QT * 128, &tmQ, qk_full, 64, t * FV_QT, h, img);
      }
      tma_load_4d(Ks, &tmK, qk_full, 0, 0, h, img);
      tma_load_4d(Ks + FV_KC * 128, &tmK, qk_full, 64, 0, h, img);
      mbar_expect_tx(v_full, FV_KV_BYTES);
      tma_load_4d(Vs, &tmV, v_full, 0, 0, h, img);
      tma_load_4d(Vs + FV_KC * 128, &tmV, v_full, 64, 0, h, img);
    }
  } else if (warp == W_MMA) {
    const uint32_t q_addr = smem_u32(Qs), k_addr = smem_u32(Ks), v_addr = smem_u32(Vs);
    constexpr uint32_t idescPV = make_idesc_bf16(FV_QT, FV_HP, 0, 1);
    const uint32_t idescS = make_idesc_bf16(FV_QT, nkeys, 0, 0);
    mbar_wait(qk_full, 0);
    tc_fence_after();
    for (int t = 0; t < NT; ++t) {
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) {
        const uint32_t qo = t * FV_Q_BYTES + ((kk < 4) ? kk * 32 : FV_QT * 128);
        const uint32_t ko = (kk < 4) ? kk * 32 : FV_KC * 128;
        uint64_t da = make_smem_desc_sw128(q_addr + qo, 16, 1024);
        uint64_t db = make_smem_desc_sw128(k_addr + ko, 16, 1024);
        umma_bf16_elect(tmem_base + t * 256, da, db, idescS, kk != 0 ? 1u : 0u);
      }
      // the commit of S(NT-1) covers the MMAs of S(0) too: only then are Q and K dead (P may overwrite them)
      if (t == NT - 1)
        for (int u = 0; u < NT; ++u) umma_commit_elect(&s_full[u]);
    }
    mbar_wait(v_full, 0);
    const int ns = (nkeys + FV_KT - 1) / FV_KT;
    for (int t = 0; t < NT; ++t) {
      const uint32_t p_addr = smem_u32(Pbuf[t]);
      mbar_wait(&s_read[t], 0);  // O(t) lives in the first 128 columns of S(t)
      for (int s2 = 0; s2 < ns; ++s2) {
        mbar_wait(&p_full[t * 4 + s2], 0);
        tc_fence_after();
        const int ksteps = min(FV_KT, nkeys - s2 * FV_KT) / 16;
        for (int kk = 0; kk < ksteps; ++kk) {
          uint64_t da = make_smem_desc_sw128(p_addr + s2 * (FV_QT * 128) + kk * 32, 16, 1024);
          uint64_t db = make_smem_desc_sw128(v_addr + (s2 * FV_KT + kk * 16) * 128, FV_KC * 128, 1024);
          umma_bf16_elect(tmem_base + t * 256, da, db, idescPV, (s2 | kk) != 0 ? 1u : 0u);
        }
      }
      umma_commit_elect(&o_full[t]);
    }
  } else if (warp < W_TMA) {
    const int wg = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const bool bf16_softmax = a.mode == 0;
    const bool leader = (warp & 3) == 0 && lane == 0;
    const bool active = wg * FV_KT < nkeys;
    auto expv = [&](float x) { return bf16_softmax ? bf16r(__expf(bf16r(x))) : __expf(x); };
    // combine one value per warpgroup across the four of them; every (tile, quantity) has its own slot, so ONE barrier
    // per exchange is enough (nothing is ever overwritten)
    auto exchange = [&](float mine, bool is_max, int slot) -> float {
      float* st = stat + slot * (FV_WG * 128);
      st[wg * 128 + r] = mine;
      fv_softmax_bar();
      float v = st[r];
#pragma unroll
      for (int o = 1; o < FV_WG; ++o) v = is_max ? fmaxf(v, st[o * 128 + r]) : v + st[o * 128 + r];
      return v;
    };
    auto epilogue = [&](int t) {  // O(t) (fp32, TMEM) -> bf16 rows; warpgroup g stores columns [32 g, 32 g + 32)
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      const int qrow = t * FV_QT + r;
      if (wg * 32 < a.hd) {
        bf16* orow = a.O + ((long)img * a.Np + qrow) * a.ldo + (long)h * a.hd;
        const int c0 = wg * 32;
        uint32_t o[32];
        tmem_ld_32x32(tmem_base + t * 256 + lane_base + c0, o);
        tmem_ld_wait();
        if (qrow < a.Np) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if (c0 + v * 8 < a.hd) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(o[8 * v + 0]), __uint_as_float(o[8 * v + 1]));
              u.y = pack_bf16x2(__uint_as_float(o[8 * v + 2]), __uint_as_float(o[8 * v + 3]));
              u.z = pack_bf16x2(__uint_as_float(o[8 * v + 4]), __uint_as_float(o[8 * v + 5]));
              u.w = pack_bf16x2(__uint_as_float(o[8 * v + 6]), __uint_as_float(o[8 * v + 7]));
              *reinterpret_cast<uint4*>(orow + c0 + v * 8) = u;
            }
          }
        }
      }
    };
    for (int t = 0; t < NT; ++t) {
      mbar_wait(&s_full[t], 0);
      tc_fence_after();
      float sv[64];
      if (active) {
        const int k0 = wg * FV_KT;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t tt[32];
          tmem_ld_32x32(tmem_base + t * 256 + lane_base + wg * FV_KT + hf * 32, tt);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c)
            sv[hf * 32 + c] = (k0 + hf * 32 + c < a.Np) ? bf16r(__uint_as_float(tt[c])) : FV_NEG_INF;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) sv[c] = FV_NEG_INF;
      }
      tc_fence_before();
      mbar_arrive_relaxed(&s_read[t]);
      float tmax = -3.4e38f;
#pragma unroll
      for (int c = 0; c < 64; ++c) tmax = fmaxf(tmax, sv[c]);
      const float m = exchange(tmax, true, 2 * t);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        sv[c] = expv(sv[c] - m);
        sum += sv[c];
      }
      const float l = exchange(sum, false, 2 * t + 1);
      // one IEEE reciprocal per row, then a multiply per element (K1 does the same): e * (1/l) differs from e / l by at
      // most one fp32 ulp, i.e. it moves a bf16 rounding with probability ~3e-5 per element
      const float inv = 1.0f / (bf16_softmax ? bf16r(l) : l);
      if (active) {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 64; c += 2) pk[c >> 1] = pack_bf16x2(sv[c] * inv, sv[c + 1] * inv);
        uint8_t* sub = Pbuf[t] + wg * (FV_QT * 128);
        uint8_t* prow = sub + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async();
        fv_wg_bar(wg);
        if (leader) {
          if (a.write_p) {
            fv_tma_store_4d(&tmP, sub, wg * FV_KT, t * FV_QT, h, img);
            fv_store_commit();
          }
          mbar_arrive(&p_full[t * 4 + wg]);
        }
      }
      if (t == 1) epilogue(0);  // O(0) has been accumulating while tile 1 went through the softmax
    }
    epilogue(NT - 1);
    if (leader && a.write_p) fv_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_ALLOC) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int make_tmap_bf16_4d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, int64_t s1,
                      int64_t s2, int64_t s3, uint32_t box0, uint32_t box1);  // gemm.cu

}  // namespace lapb

using namespace lapb;

extern "C" int lapb200_vit_attn_fwd(const void* qkv, void* O, void* P, int64_t Ni, int64_t nh, int64_t Np, int64_t hd,
                                    int64_t mode, lapb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LAPB_REQUIRE(hd % 8 == 0 && hd > 64 && hd <= 80, "vit_attn_fwd: head_dim must be in (64, 80] and a multiple of 8 (got %ld)",
               (long)hd);
  LAPB_REQUIRE(Np >= 1 && Ni >= 1 && nh >= 1 && Ni <= 65535 && nh <= 65535, "vit_attn_fwd: bad sizes");
  LAPB_REQUIRE(!P || Np % 8 == 0, "vit_attn_fwd: storing P needs Np %% 8 == 0 (got %ld)", (long)Np);
  const int64_t W = nh * hd, ld = 3 * W;
  const bf16* base = reinterpret_cast<const bf16*>(qkv);
  CUtensorMap tmQ, tmK, tmV, tmP;
  int rc;
  // dims: (head dim, token, head, image); strides in elements: token 3W, head hd, image Np*3W
  if ((rc = make_tmap_bf16_4d(&tmQ, base, hd, Np, nh, Ni, ld, hd, Np * ld, 64, FV_QT))) return rc;
  if ((rc = make_tmap_bf16_4d(&tmK, base + W, hd, Np, nh, Ni, ld, hd, Np * ld, 64, FV_KC))) return rc;
  if ((rc = make_tmap_bf16_4d(&tmV, base + 2 * W, hd, Np, nh, Ni, ld, hd, Np * ld, 64, FV_KC))) return rc;
  if (P) {
    if ((rc = make_tmap_bf16_4d(&tmP, P, Np, Np, nh, Ni, Np, Np * Np, nh * Np * Np, 64, FV_QT))) return rc;
  } else {
    tmP = tmQ;
  }
  FvArgs a;
  a.Np = (int)Np; a.NCH = (int)((Np + FV_KC - 1) / FV_KC); a.hd = (int)hd; a.mode = (int)mode; a.write_p = P ? 1 : 0;
  a.ldo = W; a.O = reinterpret_cast<bf16*>(O);
  static bool configured = false;
  if (!configured) {
    LAPB_CUDA_OK(cudaFuncSetAttribute(fa_vit_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FV_SMEM));
    LAPB_CUDA_OK(cudaFuncSetAttribute(fa_vit_fwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FVP_SMEM));
    configured = true;
  }
  static int pair_env = -1;
  if (pair_env < 0) {
    const char* e = getenv("LAPB_VIT_PAIR");
    pair_env = (e && e[0] == '0') ? 0 : 1;  // 0: always the one-tile-per-CTA kernel
  }
  // the whole row in one key chunk AND more query tiles than two waves of CTAs: one CTA per head, both query tiles pipelined
  // (measured, tools/k2_time.py: 173 vs 209 us at 64 images; at 2 images — 64 tiles, the serving case — one tile per CTA is
  //  the faster shape, 17 vs 28 us)
  if (pair_env && Np <= FV_KC && nh * Ni * (long)cdiv(Np, FV_QT) > 2L * num_sms()) {
    dim3 grid((unsigned)nh, (unsigned)Ni);
    fa_vit_fwd_pair_kernel<<<grid, FV_THREADS, FVP_SMEM, stream>>>(tmQ, tmK, tmV, tmP, a);
    LAPB_LAUNCH_OK("vit_attn_fwd (pair)");
    return 0;
  }
  dim3 grid((unsigned)cdiv(Np, FV_QT), (unsigned)nh, (unsigned)Ni);
  fa_vit_fwd_kernel<<<grid, FV_THREADS, FV_SMEM, stream>>>(tmQ, tmK, tmV, tmP, a);
  LAPB_LAUNCH_OK("vit_attn_fwd");
  return 0;
}
