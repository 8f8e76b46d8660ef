// Persistent flow-matching denoise loop (lap.py:634-672) for one sample: ALL `num_steps` Euler steps of the action expert
// in ONE cooperative kernel launch, one CTA per SM, grid-wide barriers between the dependent phases.
//
// Why: at batch 1 the action expert sees M = action_horizon = 10 rows.  Every projection is a weight-streaming GEMV-like
// product (0.63 GB of bf16 expert weights per step), the attention is 80 query rows against a 702-key cache, and a step
// is ~150 dependent kernels — the kernel-per-op path is bound by launch/dependency latency (~9 us per kernel, 1.5 ms per
// step), not by HBM (0.1 ms per step).  Here the dependency cost is one grid barrier (~2 us) per phase, the weights of the
// next phase are prefetched into L2 while a CTA waits at the barrier, and the tiny activations never leave L2.
//
// Phases (B = 1; rows m = action token a):
//   prologue  T1  time_emb(t_s) -> swish(time_mlp_in)                         for ALL steps s at once (rows r = s)
//             T2  swish(time_mlp_out) -> cond16
//             T3  mod[s, :] = cond16 @ mod_w^T + mod_b   (all 2L+1 adaRMS modulation layers: streamed ONCE, not per step)
//   per step      XE = action_in_proj(x_t)               (every CTA redundantly, in shared memory: no barrier)
//     per layer P1  h = adaRMS(XE) ; qkv = h Wqkv^T
//               P2  RoPE + attention partials: item = (head, 64-key chunk of the prefix cache | the suffix keys):
//                   S = Q_h K^T and O_c = P V by mma.sync (all 10 rows share one pass over K/V: the single KV head is read
//                   once per head, not once per (row, head)), chunk-local softmax statistics
//               P2b combine the chunks (log-sum-exp) -> O
//               P3  XE1 = XE + gate_a * (O Wo^T)
//               P4  h = adaRMS(XE1) ; act = gelu(h Wg^T) * (h Wu^T)
//               P5  XE = XE1 + gate_f * (act Wd^T)
//     final       v = action_out_proj(adaRMS(XE)) ; x_t += dt * v     (every CTA redundantly: no barrier)
// GEMM phases reuse the skinny scheme of skinny.cu: the CTA's n8 weight tiles x 8 K-splitting warps, mma.sync m16n8k16
// with the K permutation that makes every operand load one contiguous 16 bytes.
//
// Rounding points follow the kernel-per-op path (and through it SURVEY Appendix A) except one: attention probabilities are
// rounded to bf16 relative to the chunk-local maximum instead of after the global normalisation (same 2^-9 relative
// rounding per element, different grid) — tests/test_gpu_kernels.py bounds the difference.
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"
#include <stdlib.h>
#include <string.h>
#include <algorithm>

namespace lapb {

typedef __nv_bfloat16 bf16;
#define DN_BIG_NEG (-2.3819763e38f)
constexpr int DN_THREADS = 1024;
constexpr int DN_WARPS = DN_THREADS / 32;
constexpr int DN_CK = 64;  // keys per attention chunk

__device__ __forceinline__ uint4 dn_ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 dn_ldcg(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void dn_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// L2 prefetch of a contiguous byte range (16-byte aligned, multiple of 16), 16 KB per issuing thread
__device__ __forceinline__ void dn_prefetch_l2(const void* p, long bytes, int first_thread) {
  const int t = (int)threadIdx.x - first_thread;
  if (t < 0) return;
  for (long off = (long)t * 16384; off < bytes; off += (long)(DN_THREADS - first_thread) * 16384) {
    const uint32_t n = (uint32_t)((bytes - off) < 16384 ? (bytes - off) : 16384);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(p) + off), "r"(n)
                 : "memory");
  }
}
__device__ __forceinline__ void dn_unpack8(const uint4& u, float (&v)[8]) {
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
  v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 dn_pack8(const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

// Grid-wide barrier on a monotonically increasing counter (zeroed by the host before the launch; the launch is
// cooperative, so all CTAs are co-resident).  A bounded spin turns a would-be hang into an error flag.
struct GridBarrier {
  unsigned* counter;
  unsigned* error;
  unsigned target;
  unsigned nblocks;
  // light: release-RMW arrive + relaxed polling (MEMBAR.ALL.GPU, no CCTL.IVALL).  The strong form (fence.sc + ld.acquire)
  // invalidates the SM's L1 on both sides of every barrier, which turns each register-spill reload and each cached
  // parameter read after the barrier into an L2 round trip.  The light form is sufficient HERE because every value that one
  // CTA writes and another reads inside the loop travels through L2-only accesses (cp.async.cg / ld.global.cg), never L1.
  bool light;
  // arrive() / wait() split the barrier so that loads issued between the two (weights of the next phase) are not ahead of
  // the releasing fence in thread 0's program order
  __device__ __forceinline__ void arrive() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += nblocks;
      if (light) {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
      } else {
        __threadfence();
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
      }
    }
  }
  __device__ __forceinline__ void sync() {
    arrive();
    wait();
  }
  __device__ __forceinline__ void wait() {
    if (threadIdx.x == 0) {
      unsigned v, spins = 0;
      do {
        if (light)
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        else
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if ((++spins & 1023u) == 0) {  // a barrier that cannot complete must not hang the GPU: flag it and fall through
          unsigned e;
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(error) : "memory");
          if (e != 0 || spins > (1u << 21)) {
            atomicExch(error, 1u);
            break;
          }
        }
      } while (v < target);
    }
    __syncthreads();
  }
};

// Contiguous split of `n` work units over the grid.
__device__ __forceinline__ void dn_range(int n, int& begin, int& end) {
  begin = (int)(((long)blockIdx.x * n) / gridDim.x);
  end = (int)(((long)(blockIdx.x + 1) * n) / gridDim.x);
}

// cp.async (LDGSTS, L2-only) staging of activations written by other CTAs: any number in flight, no registers held
__device__ __forceinline__ void dn_cp16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void dn_cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// rows x cols bf16 (cols % 8 == 0) from global (row stride lds) to shared (row stride ldd); caller waits + syncs
__device__ __forceinline__ void dn_stage(const bf16* src, long lds, bf16* dst, int ldd, int rows, int cols) {
  const int nvec = cols >> 3;
  for (int i = threadIdx.x; i < rows * nvec; i += DN_THREADS) {
    const int r = i / nvec, c = (i % nvec) * 8;
    dn_cp16(dst + (long)r * ldd + c, src + (long)r * lds + c);
  }
}

// ---- skinny products with 32 warps ----
// A pass multiplies the staged rows As[M<=16 x K] (shared memory) with up to four n8 weight tiles.  Warp roles:
//   tile mode   (P1, P4, modulation GEMM): warp = (tslot = warp >> 3, kpart = warp & 7): tile `tslot` of the pass, K groups
//               kpart, kpart + 8, ...           (8-way K split per tile)
//   ksplit mode (P3, P5: one tile per CTA, long K): every warp works on the same tile, K groups warp, warp + 32, ...
// Either way warp w deposits its 16x8 partial tile in red[kpart][tslot] and the epilogue sums 8 (tile mode) or 32
// (ksplit mode) partials.  The weight fragments b[U] of the first K iteration are loaded by dn_load_w BEFORE the grid
// barrier that precedes the phase, so their DRAM/L2 latency overlaps the barrier.
template <int U>
__device__ __forceinline__ void dn_load_w(uint4 (&b)[U], const bf16* wtile, long ldw, int ngroups, int kg_first,
                                          int kg_stride) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int kg = kg_first + kg_stride * u;
    b[u] = (wtile != nullptr && kg < ngroups) ? dn_ld_stream(wtile + (long)g * ldw + 8 * t + kg * 32)
                                              : make_uint4(0, 0, 0, 0);
  }
}
template <int U>
__device__ __forceinline__ void dn_mma_warp(const bf16* As, int lda, int M, int ngroups, int kg_first, int kg_stride,
                                            uint4 (&b)[U], const bf16* wtile, long ldw, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bf16* xlo = As + (long)g * lda + 8 * t;
  const bf16* xhi = As + (long)(g + 8) * lda + 8 * t;
  const bool vlo = g < M, vhi = (g + 8) < M;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int base = kg_first; base < ngroups; base += kg_stride * U) {
    if (base != kg_first) dn_load_w<U>(b, wtile, ldw, ngroups, base, kg_stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kg = base + kg_stride * u;
      if (kg < ngroups) {
        const uint4 alo = vlo ? *reinterpret_cast<const uint4*>(xlo + kg * 32) : zero;
        const uint4 ahi = vhi ? *reinterpret_cast<const uint4*>(xhi + kg * 32) : zero;
        dn_mma(acc, alo.x, ahi.x, alo.y, ahi.y, b[u].x, b[u].y);
        dn_mma(acc, alo.z, ahi.z, alo.w, ahi.w, b[u].z, b[u].w);
      }
    }
  }
  // red[kpart = warp & 7][tslot = warp >> 3][lane][4]
  *reinterpret_cast<float4*>(red + ((((warp & 7) * 4 + (warp >> 3)) * 32 + lane) << 2)) =
      make_float4(acc[0], acc[1], acc[2], acc[3]);
}
// element (m, cc) of tile slot `tslot` (sum of its 8 K parts); ksplit mode sums all four slots as well
__device__ __forceinline__ float dn_tile_val(const float* red, int tslot, int m, int cc, bool all_slots) {
  const int src_lane = (m & 7) * 4 + (cc >> 1), idx = (m >> 3) * 2 + (cc & 1);
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    if (all_slots) {
#pragma unroll
      for (int q = 0; q < 4; ++q) s += red[((w * 4 + q) * 32 + src_lane) * 4 + idx];
    } else {
      s += red[((w * 4 + tslot) * 32 + src_lane) * 4 + idx];
    }
  }
  return s;
}

// adaptive RMSNorm of rows [0, A) of xe_s (gemma.py:112-131 with cond): h = bf16( x*rstd * bf16(1+scale) + shift ); warp
// per row; scale/shift come from shared memory (mod_row = [scale | shift | gate], staged with the rows).
__device__ __forceinline__ void dn_ada_norm(const bf16* xe_s, bf16* h_s, int ldh, int A, int D1, const bf16* mod_row) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = warp; m < A; m += DN_WARPS) {
    float s2 = 0.f;
    for (int c = lane * 8; c < D1; c += 256) {
      float v[8];
      dn_unpack8(*reinterpret_cast<const uint4*>(xe_s + (long)m * D1 + c), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s2 += v[j] * v[j];
    }
    s2 = warp_sum(s2);
    const float rstd = rsqrtf(s2 / D1 + 1e-6f);
    for (int c = lane * 8; c < D1; c += 256) {
      float v[8], sc[8], sh[8], o[8];
      dn_unpack8(*reinterpret_cast<const uint4*>(xe_s + (long)m * D1 + c), v);
      dn_unpack8(*reinterpret_cast<const uint4*>(mod_row + c), sc);
      dn_unpack8(*reinterpret_cast<const uint4*>(mod_row + D1 + c), sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[j] * rstd) * bf16r(1.0f + sc[j]) + sh[j];
      *reinterpret_cast<uint4*>(h_s + (long)m * ldh + c) = dn_pack8(o);
    }
  }
}

// fp32 GEMV block used by the time MLP: out[r, n] = swish( sum_k W[n,k] * in_s[r,k] + bias[n] ), warp per column
// (~40 us per call at LAP-3B size, and 8 weight loads in flight per lane change that by 6 us only: the prologue runs each
// of its code paths once, instruction-cache cold — the 4 GB prefix pass before it leaves none of the kernel's code in L2)
__device__ __noinline__ void dn_time_mlp(const float* in_s, int R, int D1, const float* W, const float* bias, float* out_f32,
                                         bf16* out_bf16) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * DN_WARPS + (threadIdx.x >> 5), nw = gridDim.x * DN_WARPS;
  for (int n = gw; n < D1; n += nw) {
    float acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = 0.f;
    const float* w = W + (long)n * D1;
    for (int k = lane; k < D1; k += 32) {
      const float wv = w[k];
#pragma unroll
      for (int r = 0; r < 16; ++r)
        if (r < R) acc[r] += wv * in_s[r * D1 + k];
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
      const float b = bias[n];
      for (int r = 0; r < R; ++r) {
        const float z = acc[r] + b;
        const float o = z / (1.0f + expf(-z));
        if (out_f32) out_f32[(long)r * D1 + n] = o;
        if (out_bf16) out_bf16[(long)r * D1 + n] = __float2bfloat16_rn(o);
      }
    }
  }
}

// ---- shared-memory carve-up (bytes), shared by the kernel and the host launcher ----
__host__ __device__ inline size_t dn_align16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline size_t dn_attn_bytes(int HD) {
  return (size_t)16 * (HD + 8) * 2      /* q_s  */ + (size_t)16 * DN_CK * 4 /* s_s */ + (size_t)16 * (DN_CK + 8) * 2 /* p_s */ +
         (size_t)16 * HD * 4            /* ks_s */ + (size_t)3 * 16 * HD * 2 /* raw q, k, v rows */ +
         (size_t)4 * 16 * DN_CK * 4     /* sp_s */;
}
// region holding, in turn: normalised rows h (P1, P4, final), attention scratch (P2), staged O (P3), staged act (P5)
__host__ __device__ inline size_t dn_big_bytes(int D1, int HD, int OD, int F1) {
  size_t b = (size_t)16 * (D1 + 8) * 2;
  const size_t a = dn_attn_bytes(HD), o = (size_t)16 * (OD + 8) * 2, f = (size_t)16 * (F1 + 8) * 2;
  b = b > a ? b : a;
  b = b > o ? b : o;
  b = b > f ? b : f;
  return dn_align16(b);
}
__host__ __device__ inline size_t dn_smem_bytes(int D1, int HD, int OD, int F1) {
  // xe_s | big | red | x_s | rope table | mod rows (attn + ffn) | mask bits     (prologue fp32 [16][D1] aliases xe_s + big)
  return (size_t)16 * D1 * 2 + dn_big_bytes(D1, HD, OD, F1) + (size_t)8 * 4 * 32 * 4 * 4 + (size_t)16 * 32 * 4 +
         (size_t)16 * (HD / 2) * 8 + (size_t)6 * D1 * 2 + (size_t)16 * 32 * 4 + 16;
}

// `fold` (LAPB_DENOISE_FOLD=1, experimental, off by default): the chunk combine P2b is done by the LAST CTA to finish a head's
// partials (per-head arrival counters in p.sync[2 .. 2+NH)), which removes one grid barrier per layer; same arithmetic in the
// same order, so the outputs are identical to the default path.
__global__ void __launch_bounds__(DN_THREADS, 1) denoise_loop_kernel(const lapb_denoise_params_t p, const int fold) {
  extern __shared__ __align__(16) unsigned char dn_smem[];
  const int A = p.A, ad = p.ad, D1 = p.D1, NH = p.NH, HD = p.HD, F1 = p.F1, L = p.L, Pn = p.Pn, W32 = p.W32;
  const int QKV = (NH + 2) * HD, OD = NH * HD, nm3 = p.nm * 3 * D1, S = p.num_steps;
  const int ldh = D1 + 8, ldo = OD + 8, ldf = F1 + 8, ldq = HD + 8, ldp = DN_CK + 8, half = HD / 2;
  // ---- shared memory ----
  bf16* xe_s = reinterpret_cast<bf16*>(dn_smem);                                     // [16][D1] residual stream copy
  unsigned char* big = dn_smem + (size_t)16 * D1 * 2;
  bf16* h_s = reinterpret_cast<bf16*>(big);                                           // [16][D1+8] / staged O / staged act
  float* red = reinterpret_cast<float*>(big + dn_big_bytes(D1, HD, OD, F1));          // [8][4][32][4]
  float* x_s = red + 8 * 4 * 32 * 4;                                                  // [16*32] x_t
  float2* rope_s = reinterpret_cast<float2*>(x_s + 16 * 32);                          // [16][HD/2] (cos, sin)
  bf16* mod_sm = reinterpret_cast<bf16*>(rope_s + 16 * half);                         // [2][3*D1]: attn-norm, ffn-norm rows
  uint32_t* bits_s = reinterpret_cast<uint32_t*>(mod_sm + 6 * D1);                    // [16][32]
  float* te_s = reinterpret_cast<float*>(dn_smem);                                    // prologue only
  // attention scratch (inside `big`)
  bf16* q_s = reinterpret_cast<bf16*>(big);                                           // [16][HD+8]
  float* s_s = reinterpret_cast<float*>(q_s + 16 * ldq);                              // [16][64]
  bf16* p_s = reinterpret_cast<bf16*>(s_s + 16 * DN_CK);                              // [16][72]
  float* ks_s = reinterpret_cast<float*>(p_s + 16 * ldp);                             // [16][HD]
  bf16* qraw = reinterpret_cast<bf16*>(ks_s + 16 * HD);                               // [16][HD] x3
  bf16* kraw = qraw + 16 * HD;
  bf16* vraw = kraw + 16 * HD;
  float* sp_s = reinterpret_cast<float*>(vraw + 16 * HD);                             // [4][16][64] S partials (K quarters)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int tslot = warp >> 3, kpart = warp & 7;
  GridBarrier bar{p.sync, p.sync + 1, 0u, gridDim.x, false};
  __shared__ int fold_last;
  // optional phase profile (thread 0 of EVERY CTA; row blockIdx.x of p.prof[grid][32]): nanoseconds per phase slot,
  // accumulated in shared memory (a global read-modify-write per tick would stall the in-order warp ~0.5 us each time)
  __shared__ unsigned long long prof_s[32];
  unsigned long long prof_last = 0;
  const bool prof_on = p.prof != nullptr && threadIdx.x == 0;
  if (prof_on)
    for (int i = 0; i < 32; ++i) prof_s[i] = 0;
  auto tick = [&](int slot) {
    if (prof_on) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (slot >= 0) prof_s[slot] += now - prof_last;
      prof_last = now;
    }
  };
  tick(-1);
  const bool split = (p.flags & 2) == 0, spread = (p.flags & 4) != 0;  // flags 0 = measured best
  const bf16* mod = reinterpret_cast<const bf16*>(p.mod);
  bf16* XE = reinterpret_cast<bf16*>(p.XE);
  bf16* XE1 = reinterpret_cast<bf16*>(p.XE1);
  bf16* qkv = reinterpret_cast<bf16*>(p.qkv);
  bf16* Obuf = reinterpret_cast<bf16*>(p.O);
  bf16* act = reinterpret_cast<bf16*>(p.act);
  const int NCHP = (Pn + DN_CK - 1) / DN_CK, NCH = NCHP + 1;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  const int ng1 = D1 >> 5, ngo = OD >> 5, ngf = F1 >> 5;

  // this CTA's share of every phase (contiguous n8 tiles / GeGLU pairs / attention item)
  int q_tb, q_te, o_tb, o_te, f_pb, f_pe;
  dn_range(QKV / 8, q_tb, q_te);
  dn_range(D1 / 8, o_tb, o_te);
  dn_range(F1 / 8, f_pb, f_pe);
  // this warp's weight tile in a tile-mode pass
  auto qkv_tile = [&](const bf16* W, int t0) { return (t0 + tslot < q_te) ? W + (long)(t0 + tslot) * 8 * D1 : nullptr; };
  auto gu_tile = [&](const bf16* W, int p0) {  // slots: gate p0, up p0, gate p0+1, up p0+1
    const int pr = p0 + (tslot >> 1);
    return (pr < f_pe) ? W + ((long)(tslot & 1) * F1 + (long)pr * 8) * D1 : nullptr;
  };

  // =========================== prologue: time conditioning of every step ===========================
  {
    // T1: time_emb (pi0.py:47-63) for all steps, then s1 = swish(time_mlp_in(time_emb))
    const int halfw = D1 / 2;
    for (int i = threadIdx.x; i < S * halfw; i += DN_THREADS) {
      const int r = i / halfw, c = i % halfw;
      const float fraction = (halfw > 1) ? (float)c / (float)(halfw - 1) : 0.f;
      const float period = 4e-3f * powf(4.0f / 4e-3f, fraction);
      const float inp = p.times[r] * (1.0f / period * 2.0f * 3.14159265358979323846f);
      float sn, cs;
      sincosf(inp, &sn, &cs);
      te_s[r * D1 + c] = sn;
      te_s[r * D1 + halfw + c] = cs;
    }
    __syncthreads();
    dn_time_mlp(te_s, S, D1, p.tin_w, p.tin_b, p.s1, nullptr);
    bar.sync();
    // T2: cond = swish(time_mlp_out(s1)) -> cond16
    for (int i = threadIdx.x; i < S * D1; i += DN_THREADS) te_s[i] = __ldcg(p.s1 + i);
    __syncthreads();
    dn_time_mlp(te_s, S, D1, p.tout_w, p.tout_b, nullptr, reinterpret_cast<bf16*>(p.cond16));
    bar.sync();
    // T3: mod = cond16 @ mod_w^T + mod_b  (rows = steps)
    dn_stage(reinterpret_cast<const bf16*>(p.cond16), D1, h_s, ldh, S, D1);
    int tb, te;
    dn_range(nm3 / 8, tb, te);
    const bf16* mw = reinterpret_cast<const bf16*>(p.mod_w);
    auto mod_tile = [&](int t0) { return (t0 + tslot < te) ? mw + (long)(t0 + tslot) * 8 * D1 : nullptr; };
    uint4 b[4];
    const bf16* wt = mod_tile(tb);
    dn_load_w<4>(b, wt, D1, ng1, kpart, 8);
    dn_cp_wait_all();
    __syncthreads();
    for (int t0 = tb; t0 < te; t0 += 4) {
      dn_mma_warp<4>(h_s, ldh, S, ng1, kpart, 8, b, wt, D1, red);
      __syncthreads();
      wt = mod_tile(t0 + 4);
      if (t0 + 4 < te) dn_load_w<4>(b, wt, D1, ng1, kpart, 8);  // next pass's weights fly during this epilogue
      if (threadIdx.x < 16 * 32) {
        const int e = threadIdx.x, m = e >> 5, c = e & 31, tile = c >> 3, cc = c & 7;
        if (m < S && t0 + tile < te) {
          const int n = (t0 + tile) * 8 + cc;
          const float v = bf16r(dn_tile_val(red, tile, m, cc, false)) + bf16r(p.mod_b[n]);
          reinterpret_cast<bf16*>(p.mod)[(long)m * nm3 + n] = __float2bfloat16_rn(v);
        }
      }
      __syncthreads();
    }
    // x_t <- noise; (cos, sin) of the suffix positions and the mask words (constant over steps and layers)
    for (int i = threadIdx.x; i < A * ad; i += DN_THREADS) x_s[i] = p.x[i];
    for (int i = threadIdx.x; i < A * half; i += DN_THREADS) {
      const int m = i / half, d = i % half;
      float sn, cs;
      sincosf((float)p.pos[m] / p.timescale[d], &sn, &cs);
      rope_s[i] = make_float2(cs, sn);
    }
    for (int i = threadIdx.x; i < A * W32; i += DN_THREADS) bits_s[(i / W32) * 32 + (i % W32)] = p.bits[i];
    bar.sync();
    tick(0);
  }

  // weights of the first P1 (layer 0): in flight while action_in_proj is computed
  uint4 w1[4];
  const bf16* wt1 = qkv_tile(reinterpret_cast<const bf16*>(p.qkv_w), q_tb);
  dn_load_w<4>(w1, wt1, D1, ng1, kpart, 8);

  // =========================== Euler loop ===========================
  for (int step = 0; step < S; ++step) {
    const bf16* mod_s = mod + (long)step * nm3;
    // XE = bf16(action_in_proj(x_t)) (pi0.py:159), every CTA holds the full copy
    for (int i = threadIdx.x; i < A * D1; i += DN_THREADS) {
      const int m = i / D1, n = i % D1;
      float v = 0.f;
      for (int j = 0; j < ad; ++j) v += x_s[m * ad + j] * __ldg(p.ain_w + (long)n * ad + j);
      xe_s[i] = __float2bfloat16_rn(v + __ldg(p.ain_b + n));
    }
    __syncthreads();
    tick(1);

    for (int l = 0; l < L; ++l) {
      const bf16* Wqkv = reinterpret_cast<const bf16*>(p.qkv_w) + (long)l * p.qkv_ls;
      const bf16* Wo = reinterpret_cast<const bf16*>(p.o_w) + (long)l * p.o_ls;
      const bf16* Wgu = reinterpret_cast<const bf16*>(p.gu_w) + (long)l * p.gu_ls;
      const bf16* Wd = reinterpret_cast<const bf16*>(p.down_w) + (long)l * p.down_ls;
      const bf16* Kc = reinterpret_cast<const bf16*>(p.Kc) + (long)l * p.kc_ls;
      const bf16* VcT = reinterpret_cast<const bf16*>(p.VcT) + (long)l * p.vct_ls;
      const bf16* mod_a = mod_sm;           // [scale | shift | gate] of the attention norm
      const bf16* mod_f = mod_sm + 3 * D1;  // ... of the ffn norm

      // Optional L2 prefetch of this CTA's weight slices of the NEXT layer (next step's layer 0 after the last one).  flags
      // bit 0: all four matrices at the top of P1 (one 34.6 MB burst per layer: measured 6 % slower than no prefetch, the small
      // dependent loads queue behind it); bit 2 (`spread`): each phase prefetches its own matrix of the next layer once its
      // staging loads have landed (neutral).  Both off by default (profiles/r02_denoise_loop.md).
      const int ln = (l + 1 < L) ? l + 1 : 0;
      const bool have_next = (l + 1 < L || step + 1 < S);
      const bf16* nq = reinterpret_cast<const bf16*>(p.qkv_w) + (long)ln * p.qkv_ls;
      const bf16* no = reinterpret_cast<const bf16*>(p.o_w) + (long)ln * p.o_ls;
      const bf16* ng = reinterpret_cast<const bf16*>(p.gu_w) + (long)ln * p.gu_ls;
      const bf16* nd = reinterpret_cast<const bf16*>(p.down_w) + (long)ln * p.down_ls;

      // ---------------- P1: h = adaRMS(XE); qkv = h Wqkv^T ----------------
      {
        if (have_next && (p.flags & 1)) {
          if (q_te > q_tb) dn_prefetch_l2(nq + (long)q_tb * 8 * D1, (long)(q_te - q_tb) * 8 * D1 * 2, 0);
          if (o_te > o_tb) dn_prefetch_l2(no + (long)o_tb * 8 * OD, (long)(o_te - o_tb) * 8 * OD * 2, 8);
          if (f_pe > f_pb) {
            dn_prefetch_l2(ng + (long)f_pb * 8 * D1, (long)(f_pe - f_pb) * 8 * D1 * 2, 16);
            dn_prefetch_l2(ng + ((long)F1 + (long)f_pb * 8) * D1, (long)(f_pe - f_pb) * 8 * D1 * 2, 24);
          }
          if (o_te > o_tb) dn_prefetch_l2(nd + (long)o_tb * 8 * F1, (long)(o_te - o_tb) * 8 * F1 * 2, 32);
        }
      }
      if (l > 0) dn_stage(XE, D1, xe_s, D1, A, D1);
      dn_stage(mod_s + (long)(2 * l) * 3 * D1, 0, mod_sm, 0, 1, 6 * D1);
      dn_cp_wait_all();
      __syncthreads();
      if (spread && have_next && q_te > q_tb) dn_prefetch_l2(nq + (long)q_tb * 8 * D1, (long)(q_te - q_tb) * 8 * D1 * 2, 0);
      tick(16);
      dn_ada_norm(xe_s, h_s, ldh, A, D1, mod_a);
      __syncthreads();
      tick(17);
      for (int t0 = q_tb; t0 < q_te; t0 += 4) {
        dn_mma_warp<4>(h_s, ldh, A, ng1, kpart, 8, w1, wt1, D1, red);
        __syncthreads();
        tick(18);
        wt1 = qkv_tile(Wqkv, t0 + 4);
        if (t0 + 4 < q_te) dn_load_w<4>(w1, wt1, D1, ng1, kpart, 8);
        if (threadIdx.x < 16 * 32) {
          const int e = threadIdx.x, m = e >> 5, c = e & 31, tile = c >> 3, cc = c & 7;
          if (m < A && t0 + tile < q_te)
            qkv[(long)m * QKV + (t0 + tile) * 8 + cc] = __float2bfloat16_rn(dn_tile_val(red, tile, m, cc, false));
        }
        __syncthreads();
      }
      // this CTA's attention item: K / V^T fragments of its key chunk do not depend on this step -> load them now
      // (S tile: warps 0..7, one 8-key tile each; P V: one n8 tile of the head dims per warp)
      tick(19);
      const int item0 = blockIdx.x;
      const bool item0_prefix = item0 < NH * NCH && (item0 % NCH) < NCHP;
      uint4 kf[2], vf[2];
      auto load_kv = [&](int key0) {
        // S tile: warp = (key tile kpart, K quarter tslot): groups tslot, tslot + 4 of the head dim
        const bf16* kr = Kc + (long)(key0 + 8 * kpart + g) * HD + 8 * t4;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int kg = tslot + 4 * u;
          kf[u] = (kg < HD / 32) ? *reinterpret_cast<const uint4*>(kr + kg * 32) : zero4;
        }
#pragma unroll
        for (int kg = 0; kg < 2; ++kg)
          vf[kg] = (warp < HD / 8)
                       ? *reinterpret_cast<const uint4*>(VcT + (long)(warp * 8 + g) * p.TpadK + key0 + 8 * t4 + kg * 32)
                       : zero4;
      };
      if (split) bar.arrive();
      if (item0_prefix) load_kv((item0 % NCH) * DN_CK);
      tick(20);
      tick(2);
      if (split) bar.wait(); else bar.sync();
      tick(3);
      if (spread && have_next && o_te > o_tb) dn_prefetch_l2(no + (long)o_tb * 8 * OD, (long)(o_te - o_tb) * 8 * OD * 2, 0);

      // ---------------- P2: attention partials, item = (head, key chunk) ----------------
      for (int item = blockIdx.x; item < NH * NCH; item += gridDim.x) {
        const int h = item / NCH, c = item % NCH;
        const bool prefix = c < NCHP;
        __syncthreads();
        dn_stage(qkv + h * HD, QKV, qraw, HD, A, HD);
        if (!prefix) {
          dn_stage(qkv + NH * HD, QKV, kraw, HD, A, HD);
          dn_stage(qkv + (NH + 1) * HD, QKV, vraw, HD, A, HD);
        }
        dn_cp_wait_all();
        __syncthreads();
        tick(24);
        // q_s <- bf16( bf16(rope(q_h)) * hd^-0.5 ) (gemma.py:215-218, 548-564); suffix item: ks_s <- bf16(rope(k))
        for (int i = threadIdx.x; i < A * half; i += DN_THREADS) {
          const int m = i / half, d = i % half;
          const float cs = rope_s[i].x, sn = rope_s[i].y;
          const float x1 = __bfloat162float(qraw[m * HD + d]), x2 = __bfloat162float(qraw[m * HD + half + d]);
          q_s[m * ldq + d] = __float2bfloat16_rn(bf16r(x1 * cs - x2 * sn) * p.qscale);
          q_s[m * ldq + half + d] = __float2bfloat16_rn(bf16r(x2 * cs + x1 * sn) * p.qscale);
          if (!prefix) {
            const float k1 = __bfloat162float(kraw[m * HD + d]), k2 = __bfloat162float(kraw[m * HD + half + d]);
            ks_s[m * HD + d] = bf16r(k1 * cs - k2 * sn);
            ks_s[m * HD + half + d] = bf16r(k2 * cs + k1 * sn);
          }
        }
        __syncthreads();
        tick(25);
        float* po = p.part_o + (long)item * 16 * HD;
        float* pml = p.part_ml + (long)item * 16 * 2;
        if (prefix) {
          // ---- prefix chunk: keys [key0, key0 + 64) of the cache, tensor cores ----
          const int key0 = c * DN_CK;
          if (item != item0) load_kv(key0);  // (only when there are more items than CTAs)
          {  // S partial tile: warp (kpart, tslot) -> keys key0 + 8*kpart .. + 8, head-dim groups tslot, tslot + 4
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const bool vlo = g < A, vhi = (g + 8) < A;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int kg = tslot + 4 * u;
              if (kg < HD / 32) {
                const uint4 alo = vlo ? *reinterpret_cast<const uint4*>(q_s + g * ldq + kg * 32 + 8 * t4) : zero4;
                const uint4 ahi = vhi ? *reinterpret_cast<const uint4*>(q_s + (g + 8) * ldq + kg * 32 + 8 * t4) : zero4;
                dn_mma(acc, alo.x, ahi.x, alo.y, ahi.y, kf[u].x, kf[u].y);
                dn_mma(acc, alo.z, ahi.z, alo.w, ahi.w, kf[u].z, kf[u].w);
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int m = g + (j >> 1) * 8, kk = 8 * kpart + 2 * t4 + (j & 1);
              sp_s[(tslot * 16 + m) * DN_CK + kk] = acc[j];
            }
          }
          __syncthreads();
          tick(26);
          // chunk-local softmax: warp m -> row m; lane -> keys lane, lane+32 (sum of the 4 K-quarter partials, mask)
          if (warp < A) {
            const int m = warp;
            float v0 = 0.f, v1 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              v0 += sp_s[(q * 16 + m) * DN_CK + lane];
              v1 += sp_s[(q * 16 + m) * DN_CK + 32 + lane];
            }
            const int k0 = key0 + lane, k1 = key0 + 32 + lane;
            const bool ok0 = k0 < Pn && ((bits_s[m * 32 + (k0 >> 5)] >> (k0 & 31)) & 1u);
            const bool ok1 = k1 < Pn && ((bits_s[m * 32 + (k1 >> 5)] >> (k1 & 31)) & 1u);
            v0 = ok0 ? v0 : DN_BIG_NEG;
            v1 = ok1 ? v1 : DN_BIG_NEG;
            const float mx = warp_max(fmaxf(v0, v1));
            const float e0 = __expf(v0 - mx), e1 = __expf(v1 - mx);
            const float sum = warp_sum(e0 + e1);
            p_s[m * ldp + lane] = __float2bfloat16_rn(e0);
            p_s[m * ldp + 32 + lane] = __float2bfloat16_rn(e1);
            if (lane == 0) {
              pml[m * 2] = mx;
              pml[m * 2 + 1] = sum;
            }
          }
          __syncthreads();
          tick(27);
          // O_c = P V : warp w -> n8 tile w of the head dims; K = 64 keys (2 groups)
          if (warp < HD / 8) {
            const bool vlo = g < A, vhi = (g + 8) < A;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kg = 0; kg < 2; ++kg) {
              const uint4 alo = vlo ? *reinterpret_cast<const uint4*>(p_s + g * ldp + kg * 32 + 8 * t4) : zero4;
              const uint4 ahi = vhi ? *reinterpret_cast<const uint4*>(p_s + (g + 8) * ldp + kg * 32 + 8 * t4) : zero4;
              dn_mma(acc, alo.x, ahi.x, alo.y, ahi.y, vf[kg].x, vf[kg].y);
              dn_mma(acc, alo.z, ahi.z, alo.w, ahi.w, vf[kg].z, vf[kg].w);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int m = g + (j >> 1) * 8, d = warp * 8 + 2 * t4 + (j & 1);
              if (m < A) po[m * HD + d] = acc[j];
            }
          }
        } else {
          // ---- suffix keys (this step's own A tokens): CUDA cores on shared memory ----
          for (int pr = warp; pr < A * A; pr += DN_WARPS) {  // logits: warp per (query a, key a2)
            const int a = pr / A, a2 = pr % A;
            float sacc = 0.f;
            for (int d = lane; d < HD; d += 32) sacc += __bfloat162float(q_s[a * ldq + d]) * ks_s[a2 * HD + d];
            sacc = warp_sum(sacc);
            if (lane == 0) {
              const int key = Pn + a2;
              const bool ok = (bits_s[a * 32 + (key >> 5)] >> (key & 31)) & 1u;
              s_s[a * DN_CK + a2] = ok ? sacc : DN_BIG_NEG;
            }
          }
          __syncthreads();
          if (warp < A) {  // warp a -> row a; lane a2 -> key a2 (A <= 16)
            const int a = warp;
            const float v = (lane < A) ? s_s[a * DN_CK + lane] : -3.4e38f;
            const float mx = warp_max(v);
            const float e = (lane < A) ? __expf(v - mx) : 0.f;
            const float sum = warp_sum(e);
            if (lane < A) s_s[a * DN_CK + lane] = bf16r(e);
            if (lane == 0) {
              pml[a * 2] = mx;
              pml[a * 2 + 1] = sum;
            }
          }
          __syncthreads();
          for (int i = threadIdx.x; i < A * HD; i += DN_THREADS) {
            const int a = i / HD, d = i % HD;
            float o = 0.f;
            for (int a2 = 0; a2 < A; ++a2) o += s_s[a * DN_CK + a2] * __bfloat162float(vraw[a2 * HD + d]);
            po[a * HD + d] = o;
          }
        }
        if (fold) {
          // last-arriver combine (the "threadfence reduction" pattern): publish this item's partials, count the arrival on
          // the head's monotonic counter, and let whichever CTA completes the head fold its NCH chunks into O
          __threadfence();
          __syncthreads();
          if (threadIdx.x == 0) {
            const unsigned old = atomicAdd(p.sync + 2 + h, 1u);
            fold_last = (old + 1u == (unsigned)(step * L + l + 1) * (unsigned)NCH) ? 1 : 0;
            __threadfence();
          }
          __syncthreads();
          if (fold_last) {
            for (int i = threadIdx.x; i < A * HD; i += DN_THREADS) {
              const int m = i / HD, d = i % HD;
              const float* ml = p.part_ml + ((long)h * NCH * 16 + m) * 2;
              const float* oc = p.part_o + ((long)h * NCH * 16 + m) * HD + d;
              float mx = -3.4e38f;
#pragma unroll 4
              for (int c2 = 0; c2 < NCH; ++c2) mx = fmaxf(mx, __ldcg(ml + (long)c2 * 32));
              float den = 0.f, num = 0.f;
#pragma unroll 4
              for (int c2 = 0; c2 < NCH; ++c2) {
                const float w = __expf(__ldcg(ml + (long)c2 * 32) - mx);
                den += w * __ldcg(ml + (long)c2 * 32 + 1);
                num += w * __ldcg(oc + (long)c2 * 16 * HD);
              }
              Obuf[(long)m * OD + h * HD + d] = __float2bfloat16_rn(num / den);
            }
          }
        }
      }
      tick(4);
      if (!fold) bar.sync();
      tick(5);

      // ---------------- P2b: combine the chunks -> O [A, NH*HD]; outputs interleaved over the grid ----------------
      const int per_cta = fold ? 0 : (A * OD + gridDim.x - 1) / gridDim.x;
      for (int i0 = threadIdx.x; i0 < per_cta; i0 += DN_THREADS) {
        const int i = blockIdx.x * per_cta + i0;
        if (i >= A * OD) break;
        const int m = i / OD, h = (i / HD) % NH, d = i % HD;
        const float* ml = p.part_ml + ((long)h * NCH * 16 + m) * 2;
        const float* oc = p.part_o + ((long)h * NCH * 16 + m) * HD + d;
        float mx = -3.4e38f;
#pragma unroll 4
        for (int c = 0; c < NCH; ++c) mx = fmaxf(mx, __ldcg(ml + (long)c * 32));
        float den = 0.f, num = 0.f;
#pragma unroll 4
        for (int c = 0; c < NCH; ++c) {
          const float w = __expf(__ldcg(ml + (long)c * 32) - mx);
          den += w * __ldcg(ml + (long)c * 32 + 1);
          num += w * __ldcg(oc + (long)c * 16 * HD);
        }
        Obuf[(long)m * OD + h * HD + d] = __float2bfloat16_rn(num / den);
      }
      tick(21);
      // P3's weights (one n8 tile of Wo per CTA, K split over all 32 warps): in flight across the barrier
      uint4 w3[2];
      const bf16* wt3 = o_te > o_tb ? Wo + (long)o_tb * 8 * OD : nullptr;
      if (split) bar.arrive();
      dn_load_w<2>(w3, wt3, OD, ngo, warp, DN_WARPS);
      tick(22);
      tick(6);
      if (split) bar.wait(); else bar.sync();
      tick(7);

      // ---------------- P3: XE1 = XE + gate_a * (O Wo^T) ----------------
      uint4 w4[4];
      const bf16* wt4 = gu_tile(Wgu, f_pb);
      if (o_te > o_tb) {
        dn_stage(Obuf, OD, h_s, ldo, A, OD);
        dn_cp_wait_all();
        __syncthreads();
        for (int t0 = o_tb; t0 < o_te; ++t0) {
          dn_mma_warp<2>(h_s, ldo, A, ngo, warp, DN_WARPS, w3, wt3, OD, red);
          __syncthreads();
          if (t0 + 1 < o_te) {
            wt3 = Wo + (long)(t0 + 1) * 8 * OD;
            dn_load_w<2>(w3, wt3, OD, ngo, warp, DN_WARPS);
          } else if (!split) {
            dn_load_w<4>(w4, wt4, D1, ng1, kpart, 8);  // P4's first pass
          }
          if (threadIdx.x < 16 * 8) {
            const int m = threadIdx.x >> 3, cc = threadIdx.x & 7;
            if (m < A) {
              const int n = t0 * 8 + cc;
              const float y = bf16r(dn_tile_val(red, 0, m, cc, true));
              const float gt = __bfloat162float(mod_a[2 * D1 + n]);
              const float r = __bfloat162float(xe_s[m * D1 + n]);
              XE1[(long)m * D1 + n] = __float2bfloat16_rn(r + bf16r(y * gt));
            }
          }
          __syncthreads();
        }
      } else if (!split) {
        dn_load_w<4>(w4, wt4, D1, ng1, kpart, 8);
      }
      tick(8);
      if (split) {
        bar.arrive();
        dn_load_w<4>(w4, wt4, D1, ng1, kpart, 8);
        bar.wait();
      } else {
        bar.sync();
      }
      tick(9);

      // ---------------- P4: h = adaRMS(XE1); act = gelu(h Wg^T) * (h Wu^T) ----------------
      dn_stage(XE1, D1, xe_s, D1, A, D1);
      dn_cp_wait_all();
      __syncthreads();
      if (spread && have_next && f_pe > f_pb) {
        dn_prefetch_l2(ng + (long)f_pb * 8 * D1, (long)(f_pe - f_pb) * 8 * D1 * 2, 0);
        dn_prefetch_l2(ng + ((long)F1 + (long)f_pb * 8) * D1, (long)(f_pe - f_pb) * 8 * D1 * 2, 8);
      }
      tick(28);
      dn_ada_norm(xe_s, h_s, ldh, A, D1, mod_f);
      __syncthreads();
      tick(29);
      for (int p0 = f_pb; p0 < f_pe; p0 += 2) {
        dn_mma_warp<4>(h_s, ldh, A, ng1, kpart, 8, w4, wt4, D1, red);
        __syncthreads();
        wt4 = gu_tile(Wgu, p0 + 2);
        if (p0 + 2 < f_pe) dn_load_w<4>(w4, wt4, D1, ng1, kpart, 8);
        if (threadIdx.x < 256) {
          const int e = threadIdx.x;  // 16 rows x 2 pairs x 8 columns = 256 outputs
          const int m = e >> 4, pi = (e >> 3) & 1, cc = e & 7;
          if (m < A && p0 + pi < f_pe) {
            const float gv = bf16r(dn_tile_val(red, 2 * pi, m, cc, false));
            const float uv = bf16r(dn_tile_val(red, 2 * pi + 1, m, cc, false));
            act[(long)m * F1 + (p0 + pi) * 8 + cc] = __float2bfloat16_rn(bf16r(gelu_tanh(gv)) * uv);
          }
        }
        __syncthreads();
      }
      tick(30);
      // P5's tile of Wd (K split over all 32 warps): in flight across the barrier
      uint4 w5[4];
      const bf16* wt5 = o_te > o_tb ? Wd + (long)o_tb * 8 * F1 : nullptr;
      if (split) bar.arrive();
      dn_load_w<4>(w5, wt5, F1, ngf, warp, DN_WARPS);
      tick(10);
      if (split) bar.wait(); else bar.sync();
      tick(11);

      // ---------------- P5: XE = XE1 + gate_f * (act Wd^T) ----------------
      if (o_te > o_tb) {
        dn_stage(act, F1, h_s, ldf, A, F1);
        dn_cp_wait_all();
        __syncthreads();
        if (spread && have_next) dn_prefetch_l2(nd + (long)o_tb * 8 * F1, (long)(o_te - o_tb) * 8 * F1 * 2, 0);
        for (int t0 = o_tb; t0 < o_te; ++t0) {
          dn_mma_warp<4>(h_s, ldf, A, ngf, warp, DN_WARPS, w5, wt5, F1, red);
          __syncthreads();
          if (t0 + 1 < o_te) {
            wt5 = Wd + (long)(t0 + 1) * 8 * F1;
            dn_load_w<4>(w5, wt5, F1, ngf, warp, DN_WARPS);
          }
          if (threadIdx.x < 16 * 8) {
            const int m = threadIdx.x >> 3, cc = threadIdx.x & 7;
            if (m < A) {
              const int n = t0 * 8 + cc;
              const float y = bf16r(dn_tile_val(red, 0, m, cc, true));
              const float gt = __bfloat162float(mod_f[2 * D1 + n]);
              const float r = __bfloat162float(xe_s[m * D1 + n]);
              XE[(long)m * D1 + n] = __float2bfloat16_rn(r + bf16r(y * gt));
            }
          }
          __syncthreads();
        }
      }
      // next layer's qkv tiles: in flight across the barrier (the next step's layer 0 is loaded in the final phase)
      if (split) bar.arrive();
      if (l + 1 < L) {
        wt1 = qkv_tile(Wqkv + p.qkv_ls, q_tb);
        dn_load_w<4>(w1, wt1, D1, ng1, kpart, 8);
      }
      tick(12);
      if (split) bar.wait(); else bar.sync();
      tick(13);
    }

    // ---------------- final: v = action_out_proj(adaRMS(XE)); x += dt * v (lap.py:665-667) ----------------
    if (step + 1 < S) {
      wt1 = qkv_tile(reinterpret_cast<const bf16*>(p.qkv_w), q_tb);
      dn_load_w<4>(w1, wt1, D1, ng1, kpart, 8);
    }
    dn_stage(XE, D1, xe_s, D1, A, D1);
    dn_stage(mod_s + (long)(p.nm - 1) * 3 * D1, 0, mod_sm, 0, 1, 2 * D1);
    dn_cp_wait_all();
    __syncthreads();
    dn_ada_norm(xe_s, h_s, ldh, A, D1, mod_sm);
    __syncthreads();
    for (int o = warp; o < A * ad; o += DN_WARPS) {
      const int m = o / ad, j = o % ad;
      float acc = 0.f;
#pragma unroll 8
      for (int k = lane; k < D1; k += 32) acc += __ldg(p.aout_w + (long)j * D1 + k) * __bfloat162float(h_s[m * ldh + k]);
      acc = warp_sum(acc);
      if (lane == 0) x_s[o] += p.dt * (acc + __ldg(p.aout_b + j));
    }
    __syncthreads();
    tick(14);
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < A * ad; i += DN_THREADS) p.x[i] = x_s[i];
  if (prof_on)
    for (int i = 0; i < 32; ++i) p.prof[(long)blockIdx.x * 32 + i] += prof_s[i];
}

// ======================================================================================================================
// K10 v2 — the same grid-wide Euler loop, restructured after the per-CTA phase profile of round 2
// (profiles/r02_denoise_loop.md §4).  What the profile showed about the kernel above:
//   * 1024 threads cap the kernel at 64 registers, so the weight fragments "preloaded into registers across the barrier"
//     were spilled: the thread waited for the HBM load to land in order to store it to local memory — the latency the
//     preload was meant to hide sat on the critical path of every phase (1.6-3.5 us at five sites per layer);
//   * fence.sc + ld.acquire invalidate the L1 on both sides of every barrier, so each spill reload / parameter read after
//     a barrier was an L2 round trip;
//   * the chunk combine walked 12 chunks x 3 dependent L2 loads per output (3 us); the adaRMS norms used 10 of 32 warps.
// Changes (arithmetic and rounding points unchanged, only fp32 summation orders of the norm statistic / chunk combine /
// action_out_proj differ):
//   * weights and K/V fragments travel global -> shared with cp.async into per-THREAD private 16-byte slots (`wbuf`:
//     [warp][4][lane]): nothing is held in registers, the copy is issued right after the slot's last read and waited for
//     with the staging loads of the phase that consumes it; both gate/up passes of P4 are in flight at once (second
//     buffer in the idle part of `big`);
//   * light grid barrier (red.release + relaxed polling: no L1 invalidation), arrive/wait split, copies issued between;
//   * chunk combine: 4 lanes per 4 outputs, chunks strided over the lanes (one L2 latency, 5 live warps); the suffix
//     chunk of a head (CUDA cores, the straggler of P2) is split by query rows over the CTAs without a prefix chunk;
//   * adaRMS norm on 2 warps per row; action_in_proj with a thread per column; activations staged with A rows
//     (not 16) and 64-byte row padding (conflict-free A-fragment reads).
// p.flags (experiments): bit 1 strong barrier (fence.sc + ld.acquire).
constexpr int DN2_PAD = 32;                                    // bf16 elements: rows 64 B apart modulo 128 B
constexpr size_t DN2_WBUF = (size_t)DN_WARPS * 4 * 32 * 16;    // 64 KB: [warp][slot 0..3][lane] uint4

__host__ __device__ inline size_t dn2_h_bytes(int rows, int D1) { return dn_align16((size_t)rows * (D1 + DN2_PAD) * 2); }
// region holding, in turn: normalised rows h (+ the second weight buffer behind them: P4, modulation GEMM), attention
// scratch (P2), staged O (P3), staged act (P5)
__host__ __device__ inline size_t dn2_big_bytes(int A, int D1, int HD, int OD, int F1) {
  size_t b = dn2_h_bytes(A, D1) + DN2_WBUF;
  const size_t a = dn_attn_bytes(HD), o = (size_t)A * (OD + DN2_PAD) * 2, f = (size_t)A * (F1 + DN2_PAD) * 2;
  b = b > a ? b : a;
  b = b > o ? b : o;
  b = b > f ? b : f;
  return dn_align16(b);
}
__host__ __device__ inline size_t dn2_xe_bytes(int A, int D1) { return dn_align16((size_t)A * D1 * 2); }  // [A][D1] bf16
// xe_s | big; the prologue lays its own buffers over both: fp32 [S][D1] time embedding (T1, T2), then the staged cond16
// rows [S][D1+PAD] followed by the second weight buffer (T3)
__host__ __device__ inline size_t dn2_act_bytes(int A, int S, int D1, int HD, int OD, int F1) {
  size_t act = dn2_xe_bytes(A, D1) + dn2_big_bytes(A, D1, HD, OD, F1);
  const size_t pro = dn_align16((size_t)S * D1 * 4), t3 = dn2_h_bytes(S, D1) + DN2_WBUF;
  act = act > pro ? act : pro;
  act = act > t3 ? act : t3;
  return act;
}
__host__ __device__ inline size_t dn2_smem_bytes(int A, int S, int D1, int HD, int OD, int F1) {
  // xe_s | big | wbuf | red | x_s | rope table | mod rows (attn + ffn) | mask bits
  const size_t act = dn2_act_bytes(A, S, D1, HD, OD, F1);
  return act + DN2_WBUF + (size_t)8 * 4 * 32 * 4 * 4 + (size_t)16 * 32 * 4 + (size_t)A * (HD / 2) * 8 + (size_t)6 * D1 * 2 +
         (size_t)16 * 32 * 4 + 16;
}

// this thread's U fragments of one n8 weight tile (K groups kg_first + kg_stride * u) -> its private slots wme[u * 32].
// wtile == nullptr: the warp's tile slot lies beyond the CTA's share — nothing is copied, dn2_mma_warp skips the warp and
// the epilogues never read that slot of `red`.
template <int U>
__device__ __forceinline__ void dn2_issue_w(uint4* wme, const bf16* wtile, long ldw, int ngroups, int kg_first,
                                            int kg_stride) {
  if (wtile == nullptr) return;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int kg = kg_first + kg_stride * u;
    if (kg < ngroups) dn_cp16(wme + u * 32, wtile + (long)g * ldw + 8 * t + kg * 32);
  }
}
// as dn_mma_warp, B fragments from the thread's slots (the caller has waited for the copies of the first K batch)
template <int U>
__device__ __forceinline__ void dn2_mma_warp(const bf16* As, int lda, int M, int ngroups, int kg_first, int kg_stride,
                                             uint4* wme, const bf16* wtile, long ldw, float* red) {
  if (wtile == nullptr) return;  // (warp-uniform) no tile in this slot: saves its share of the shared-memory bandwidth
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bf16* xlo = As + (long)g * lda + 8 * t;
  const bf16* xhi = As + (long)(g + 8) * lda + 8 * t;
  const bool vlo = g < M, vhi = (g + 8) < M;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int base = kg_first; base < ngroups; base += kg_stride * U) {
    if (base != kg_first) {  // K longer than one batch (not the case at LAP-3B sizes): refill the slots and wait
      dn2_issue_w<U>(wme, wtile, ldw, ngroups, base, kg_stride);
      dn_cp_wait_all();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kg = base + kg_stride * u;
      if (kg < ngroups) {
        const uint4 b = wme[u * 32];
        const uint4 alo = vlo ? *reinterpret_cast<const uint4*>(xlo + kg * 32) : zero;
        const uint4 ahi = vhi ? *reinterpret_cast<const uint4*>(xhi + kg * 32) : zero;
        dn_mma(acc, alo.x, ahi.x, alo.y, ahi.y, b.x, b.y);
        dn_mma(acc, alo.z, ahi.z, alo.w, ahi.w, b.z, b.w);
      }
    }
  }
  *reinterpret_cast<float4*>(red + ((((warp & 7) * 4 + (warp >> 3)) * 32 + lane) << 2)) =
      make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// adaptive RMSNorm (see dn_ada_norm) on 2 warps per row: warp = (row m = warp >> 1, column half warp & 1); `scratch` holds
// the 2A partial sums of squares.  Contains one __syncthreads().
__device__ __forceinline__ void dn2_ada_norm(const bf16* xe_s, bf16* h_s, int ldh, int A, int D1, const bf16* mod_row,
                                             float* scratch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = warp >> 1, c0 = (warp & 1) * (D1 >> 1), c1 = c0 + (D1 >> 1);
  float s2 = 0.f;
  if (m < A) {
    for (int c = c0 + lane * 8; c < c1; c += 256) {
      float v[8];
      dn_unpack8(*reinterpret_cast<const uint4*>(xe_s + (long)m * D1 + c), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s2 += v[j] * v[j];
    }
  }
  s2 = warp_sum(s2);
  if (lane == 0) scratch[warp] = s2;
  __syncthreads();
  if (m < A) {
    const float rstd = rsqrtf((scratch[2 * m] + scratch[2 * m + 1]) / D1 + 1e-6f);
    for (int c = c0 + lane * 8; c < c1; c += 256) {
      float v[8], sc[8], sh[8], o[8];
      dn_unpack8(*reinterpret_cast<const uint4*>(xe_s + (long)m * D1 + c), v);
      dn_unpack8(*reinterpret_cast<const uint4*>(mod_row + c), sc);
      dn_unpack8(*reinterpret_cast<const uint4*>(mod_row + D1 + c), sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[j] * rstd) * bf16r(1.0f + sc[j]) + sh[j];
      *reinterpret_cast<uint4*>(h_s + (long)m * ldh + c) = dn_pack8(o);
    }
  }
}

template <bool PROF>
__global__ void __launch_bounds__(DN_THREADS, 1) denoise_loop2_kernel(const lapb_denoise_params_t p) {
  extern __shared__ __align__(16) unsigned char dn_smem[];
  const int A = p.A, ad = p.ad, D1 = p.D1, NH = p.NH, HD = p.HD, F1 = p.F1, L = p.L, Pn = p.Pn, W32 = p.W32;
  const int QKV = (NH + 2) * HD, OD = NH * HD, nm3 = p.nm * 3 * D1, S = p.num_steps;
  const int ldh = D1 + DN2_PAD, ldo = OD + DN2_PAD, ldf = F1 + DN2_PAD, ldq = HD + 8, ldp = DN_CK + 8, half = HD / 2;
  // ---- shared memory ----
  bf16* xe_s = reinterpret_cast<bf16*>(dn_smem);                                     // [A][D1] residual stream copy
  unsigned char* big = dn_smem + dn2_xe_bytes(A, D1);
  bf16* h_s = reinterpret_cast<bf16*>(big);                                           // [A][D1+PAD] / staged O / staged act
  unsigned char* wbuf_b = dn_smem + dn2_act_bytes(A, S, D1, HD, OD, F1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4* wme = reinterpret_cast<uint4*>(wbuf_b) + warp * 128 + lane;                                   // slots wme[u * 32]
  uint4* wme2 = reinterpret_cast<uint4*>(big + dn2_h_bytes(A, D1)) + warp * 128 + lane;                // second buffer
  float* red = reinterpret_cast<float*>(wbuf_b + DN2_WBUF);                           // [8][4][32][4]
  float* x_s = red + 8 * 4 * 32 * 4;                                                  // [16*32] x_t
  float2* rope_s = reinterpret_cast<float2*>(x_s + 16 * 32);                          // [A][HD/2] (cos, sin)
  bf16* mod_sm = reinterpret_cast<bf16*>(rope_s + A * half);                          // [2][3*D1]: attn-norm, ffn-norm rows
  uint32_t* bits_s = reinterpret_cast<uint32_t*>(mod_sm + 6 * D1);                    // [16][32]
  float* te_s = reinterpret_cast<float*>(dn_smem);                                    // prologue only
  // attention scratch (inside `big`)
  bf16* q_s = reinterpret_cast<bf16*>(big);                                           // [16][HD+8]
  float* s_s = reinterpret_cast<float*>(q_s + 16 * ldq);                              // [16][64]
  bf16* p_s = reinterpret_cast<bf16*>(s_s + 16 * DN_CK);                              // [16][72]
  float* ks_s = reinterpret_cast<float*>(p_s + 16 * ldp);                             // [16][HD]
  bf16* vraw = reinterpret_cast<bf16*>(ks_s + 16 * HD) + 2 * 16 * HD;                 // [16][HD] staged suffix v rows
  float* sp_s = reinterpret_cast<float*>(vraw + 16 * HD);                             // [4][16][64] S partials (K quarters)

  const int g = lane >> 2, t4 = lane & 3;
  const int tslot = warp >> 3, kpart = warp & 7;
  GridBarrier bar{p.sync, p.sync + 1, 0u, gridDim.x, (p.flags & 2) == 0};
  // (the next phase's copies are issued AFTER the arrive: issued before it, by the warps that do not hold the releasing
  //  thread, the loop is 4 % slower — the releasing MEMBAR also drains the copies in flight)
  // optional phase profile: thread 0 of every CTA, row blockIdx.x of p.prof[grid][32], accumulated in shared memory
  __shared__ unsigned long long prof_s[32];
  unsigned long long prof_last = 0;
  const bool prof_on = PROF && p.prof != nullptr && threadIdx.x == 0;
  if (PROF && prof_on)
    for (int i = 0; i < 32; ++i) prof_s[i] = 0;
  auto tick = [&](int slot) {
    if (PROF && prof_on) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (slot >= 0) prof_s[slot] += now - prof_last;
      prof_last = now;
    }
  };
  tick(-1);
  const bf16* mod = reinterpret_cast<const bf16*>(p.mod);
  bf16* XE = reinterpret_cast<bf16*>(p.XE);
  bf16* XE1 = reinterpret_cast<bf16*>(p.XE1);
  bf16* qkv = reinterpret_cast<bf16*>(p.qkv);
  bf16* Obuf = reinterpret_cast<bf16*>(p.O);
  bf16* act = reinterpret_cast<bf16*>(p.act);
  const int NCHP = (Pn + DN_CK - 1) / DN_CK, NCH = NCHP + 1;
  const int nprefix = NH * NCHP;
  int sfx_rows, sfx_blocks;  // suffix chunk of a head: sfx_blocks items of sfx_rows query rows
  {
    int r = ((int)gridDim.x > nprefix) ? ((int)gridDim.x - nprefix) / NH : 1;
    r = r < 1 ? 1 : (r > A ? A : r);
    sfx_rows = (A + r - 1) / r;
    sfx_blocks = (A + sfx_rows - 1) / sfx_rows;
  }
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  const int ng1 = D1 >> 5, ngo = OD >> 5, ngf = F1 >> 5;

  // this CTA's share of every phase (contiguous n8 tiles / GeGLU pairs / attention item)
  int q_tb, q_te, o_tb, o_te, f_pb, f_pe;
  dn_range(QKV / 8, q_tb, q_te);
  dn_range(D1 / 8, o_tb, o_te);
  dn_range(F1 / 8, f_pb, f_pe);
  // this warp's weight tile in a tile-mode pass
  auto qkv_tile = [&](const bf16* W, int t0) { return (t0 + tslot < q_te) ? W + (long)(t0 + tslot) * 8 * D1 : nullptr; };
  auto gu_tile = [&](const bf16* W, int p0) {  // slots: gate p0, up p0, gate p0+1, up p0+1
    const int pr = p0 + (tslot >> 1);
    return (pr < f_pe) ? W + ((long)(tslot & 1) * F1 + (long)pr * 8) * D1 : nullptr;
  };
  // K / V^T fragments of a 64-key prefix chunk -> slots 0,1 (S tile: warp = (key tile kpart, head-dim quarter tslot)) and
  // 2,3 (P V: warp = n8 tile of the head dims).  (cp.async with an L2 evict_last cache hint for these 13 MB faults with
  // "illegal instruction" on sm_100a under CUDA 12.9 — plain copies.)
  auto issue_kv = [&](const bf16* Kc, const bf16* VcT, int key0) {
    const bf16* kr = Kc + (long)(key0 + 8 * kpart + g) * HD + 8 * t4;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int kg = tslot + 4 * u;
      if (kg < HD / 32) dn_cp16(wme + u * 32, kr + kg * 32);
    }
#pragma unroll
    for (int kg = 0; kg < 2; ++kg) {
      if (warp < HD / 8) dn_cp16(wme + (2 + kg) * 32, VcT + (long)(warp * 8 + g) * p.TpadK + key0 + 8 * t4 + kg * 32);
    }
  };

  // =========================== prologue: time conditioning of every step ===========================
  {
    // T1: time_emb (pi0.py:47-63) for all steps, then s1 = swish(time_mlp_in(time_emb))
    const int halfw = D1 / 2;
    for (int i = threadIdx.x; i < S * halfw; i += DN_THREADS) {
      const int r = i / halfw, c = i % halfw;
      const float fraction = (halfw > 1) ? (float)c / (float)(halfw - 1) : 0.f;
      const float period = 4e-3f * powf(4.0f / 4e-3f, fraction);
      const float inp = p.times[r] * (1.0f / period * 2.0f * 3.14159265358979323846f);
      float sn, cs;
      sincosf(inp, &sn, &cs);
      te_s[r * D1 + c] = sn;
      te_s[r * D1 + halfw + c] = cs;
    }
    __syncthreads();
    dn_time_mlp(te_s, S, D1, p.tin_w, p.tin_b, p.s1, nullptr);
    bar.sync();
    tick(15);
    // T2: cond = swish(time_mlp_out(s1)) -> cond16
    for (int i = threadIdx.x; i < S * D1; i += DN_THREADS) te_s[i] = __ldcg(p.s1 + i);
    __syncthreads();
    dn_time_mlp(te_s, S, D1, p.tout_w, p.tout_b, nullptr, reinterpret_cast<bf16*>(p.cond16));
    bar.sync();
    tick(23);
    // T3: mod = cond16 @ mod_w^T + mod_b  (rows = steps); passes alternate between the two weight buffers, two in flight
    bf16* h3 = reinterpret_cast<bf16*>(dn_smem);                                      // [S][D1+PAD] staged cond16
    uint4* wme3 = reinterpret_cast<uint4*>(dn_smem + dn2_h_bytes(S, D1)) + warp * 128 + lane;
    dn_stage(reinterpret_cast<const bf16*>(p.cond16), D1, h3, ldh, S, D1);
    int tb, te;
    dn_range(nm3 / 8, tb, te);
    const bf16* mw = reinterpret_cast<const bf16*>(p.mod_w);
    auto mod_tile = [&](int t0) { return (t0 + tslot < te) ? mw + (long)(t0 + tslot) * 8 * D1 : nullptr; };
    if (tb < te) dn2_issue_w<4>(wme, mod_tile(tb), D1, ng1, kpart, 8);
    if (tb + 4 < te) dn2_issue_w<4>(wme3, mod_tile(tb + 4), D1, ng1, kpart, 8);
    dn_cp_wait_all();
    __syncthreads();
    int pass = 0;
    for (int t0 = tb; t0 < te; t0 += 4, ++pass) {
      uint4* wb = (pass & 1) ? wme3 : wme;
      float bias_v = 0.f;  // this thread's output column of the pass: its bias is in flight during the product
      {
        const int e = threadIdx.x, c = e & 31, tile = c >> 3, cc = c & 7;
        if (e < 16 * 32 && (e >> 5) < S && t0 + tile < te) bias_v = __ldg(p.mod_b + (t0 + tile) * 8 + cc);
      }
      dn2_mma_warp<4>(h3, ldh, S, ng1, kpart, 8, wb, mod_tile(t0), D1, red);
      if (t0 + 8 < te) dn2_issue_w<4>(wb, mod_tile(t0 + 8), D1, ng1, kpart, 8);  // refill: consumed two passes from now
      __syncthreads();
      if (threadIdx.x < 16 * 32) {
        const int e = threadIdx.x, m = e >> 5, c = e & 31, tile = c >> 3, cc = c & 7;
        if (m < S && t0 + tile < te) {
          const int n = (t0 + tile) * 8 + cc;
          const float v = bf16r(dn_tile_val(red, tile, m, cc, false)) + bf16r(bias_v);
          reinterpret_cast<bf16*>(p.mod)[(long)m * nm3 + n] = __float2bfloat16_rn(v);
        }
      }
      // the next pass's weights were issued one pass ago; cp.async completes in order, so allow the refill just issued
      // to stay in flight: wait_group needs commit groups, so simply wait for everything except when a refill was issued
      if (t0 + 8 < te) {
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        dn_cp_wait_all();
      }
      __syncthreads();
    }
    // x_t <- noise; (cos, sin) of the suffix positions and the mask words (constant over steps and layers)
    for (int i = threadIdx.x; i < A * ad; i += DN_THREADS) x_s[i] = p.x[i];
    for (int i = threadIdx.x; i < A * half; i += DN_THREADS) {
      const int m = i / half, d = i % half;
      float sn, cs;
      sincosf((float)p.pos[m] / p.timescale[d], &sn, &cs);
      rope_s[i] = make_float2(cs, sn);
    }
    for (int i = threadIdx.x; i < A * W32; i += DN_THREADS) bits_s[(i / W32) * 32 + (i % W32)] = p.bits[i];
    bar.arrive();
    // Wqkv tiles of layer 0: in flight across the barrier and action_in_proj
    dn2_issue_w<4>(wme, qkv_tile(reinterpret_cast<const bf16*>(p.qkv_w), q_tb), D1, ng1, kpart, 8);
    bar.wait();
    tick(0);
  }

  // =========================== Euler loop ===========================
  for (int step = 0; step < S; ++step) {
    const bf16* mod_s = mod + (long)step * nm3;
    // XE = bf16(action_in_proj(x_t)) (pi0.py:159), every CTA holds the full copy; thread = output column
    for (int n = threadIdx.x; n < D1; n += DN_THREADS) {
      float acc[16];
#pragma unroll
      for (int m = 0; m < 16; ++m) acc[m] = 0.f;
      for (int j0 = 0; j0 < ad; j0 += 8) {  // (the L1 is ~6 KB next to 222 KB of shared memory: these are L2 reads)
        float wv[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) wv[jj] = (j0 + jj < ad) ? __ldg(p.ain_w + (long)n * ad + j0 + jj) : 0.f;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          if (j0 + jj < ad) {
#pragma unroll
            for (int m = 0; m < 16; ++m)
              if (m < A) acc[m] += x_s[m * ad + j0 + jj] * wv[jj];
          }
        }
      }
      const float bv = __ldg(p.ain_b + n);
#pragma unroll
      for (int m = 0; m < 16; ++m)
        if (m < A) xe_s[m * D1 + n] = __float2bfloat16_rn(acc[m] + bv);
    }
    __syncthreads();
    tick(1);

    for (int l = 0; l < L; ++l) {
      const bf16* Wqkv = reinterpret_cast<const bf16*>(p.qkv_w) + (long)l * p.qkv_ls;
      const bf16* mod_a = mod_sm;           // [scale | shift | gate] of the attention norm
      const bf16* mod_f = mod_sm + 3 * D1;  // ... of the ffn norm

      // ---------------- P1: h = adaRMS(XE); qkv = h Wqkv^T ----------------
      if (l > 0) dn_stage(XE, D1, xe_s, D1, A, D1);
      dn_stage(mod_s + (long)(2 * l) * 3 * D1, 0, mod_sm, 0, 1, 6 * D1);
      dn_cp_wait_all();  // + this CTA's Wqkv fragments, issued before the previous barrier
      __syncthreads();
      tick(16);
      dn2_ada_norm(xe_s, h_s, ldh, A, D1, mod_a, red);
      __syncthreads();
      tick(17);
      for (int t0 = q_tb; t0 < q_te; t0 += 4) {
        dn2_mma_warp<4>(h_s, ldh, A, ng1, kpart, 8, wme, qkv_tile(Wqkv, t0), D1, red);
        const bool more = t0 + 4 < q_te;
        if (more) dn2_issue_w<4>(wme, qkv_tile(Wqkv, t0 + 4), D1, ng1, kpart, 8);
        __syncthreads();
        tick(18);
        if (threadIdx.x < 16 * 32) {
          const int e = threadIdx.x, m = e >> 5, c = e & 31, tile = c >> 3, cc = c & 7;
          if (m < A && t0 + tile < q_te)
            qkv[(long)m * QKV + (t0 + tile) * 8 + cc] = __float2bfloat16_rn(dn_tile_val(red, tile, m, cc, false));
        }
        if (more) {
          dn_cp_wait_all();
          __syncthreads();
        }
      }
      tick(19);
      // this CTA's attention item: the K / V^T fragments of its key chunk do not depend on this step -> copy them now.
      // Items: (head, prefix chunk) for item < nprefix, then (head, block of query rows) of the suffix keys — the suffix
      // chunk runs on CUDA cores and would be the straggler of the phase as ONE item per head; its rows are independent
      // (per-row softmax statistics), so it is split over the CTAs the prefix chunks leave idle.
      const int item0 = blockIdx.x;
      const bool item0_prefix = item0 < nprefix;
      {
        const bf16* Kc = reinterpret_cast<const bf16*>(p.Kc) + (long)l * p.kc_ls;
        const bf16* VcT = reinterpret_cast<const bf16*>(p.VcT) + (long)l * p.vct_ls;
        bar.arrive();
        if (item0_prefix) issue_kv(Kc, VcT, (item0 % NCHP) * DN_CK);
        tick(20);
        tick(2);
        bar.wait();
        tick(3);
      }

      // ---------------- P2: attention partials, item = (head, key chunk) ----------------
      for (int item = blockIdx.x; item < nprefix + NH * sfx_blocks; item += gridDim.x) {
        const bool prefix = item < nprefix;
        const int h = prefix ? item / NCHP : (item - nprefix) / sfx_blocks, c = prefix ? item % NCHP : NCHP;
        const int r0 = prefix ? 0 : ((item - nprefix) % sfx_blocks) * sfx_rows;  // suffix item: query rows [r0, r1)
        const int r1 = prefix ? A : min(A, r0 + sfx_rows);
        __syncthreads();
        if (!prefix) {
          dn_stage(qkv + (NH + 1) * HD, QKV, vraw, HD, A, HD);
        } else if (item != item0) {  // (only when there are more items than CTAs)
          issue_kv(reinterpret_cast<const bf16*>(p.Kc) + (long)l * p.kc_ls,
                   reinterpret_cast<const bf16*>(p.VcT) + (long)l * p.vct_ls, c * DN_CK);
        }
        // q_s <- bf16( bf16(rope(q_h)) * hd^-0.5 ) (gemma.py:215-218, 548-564); suffix item: ks_s <- bf16(rope(k)).  The rows
        // come straight from L2 (two adjacent head dims per thread), RoPE applied on the way
        for (int i = threadIdx.x; i < A * (half >> 1); i += DN_THREADS) {
          const int m = i / (half >> 1), d = (i % (half >> 1)) * 2;
          const float2 c0 = rope_s[m * half + d], c1 = rope_s[m * half + d + 1];  // (cos, sin)
          const bf16* qrow = qkv + (long)m * QKV + h * HD;
          const float2 x1 = unpack_bf16x2(__ldcg(reinterpret_cast<const uint32_t*>(qrow + d)));
          const float2 x2 = unpack_bf16x2(__ldcg(reinterpret_cast<const uint32_t*>(qrow + half + d)));
          *reinterpret_cast<uint32_t*>(q_s + m * ldq + d) =
              pack_bf16x2(bf16r(x1.x * c0.x - x2.x * c0.y) * p.qscale, bf16r(x1.y * c1.x - x2.y * c1.y) * p.qscale);
          *reinterpret_cast<uint32_t*>(q_s + m * ldq + half + d) =
              pack_bf16x2(bf16r(x2.x * c0.x + x1.x * c0.y) * p.qscale, bf16r(x2.y * c1.x + x1.y * c1.y) * p.qscale);
          if (!prefix) {
            const bf16* krow = qkv + (long)m * QKV + NH * HD;
            const float2 k1 = unpack_bf16x2(__ldcg(reinterpret_cast<const uint32_t*>(krow + d)));
            const float2 k2 = unpack_bf16x2(__ldcg(reinterpret_cast<const uint32_t*>(krow + half + d)));
            ks_s[m * HD + d] = bf16r(k1.x * c0.x - k2.x * c0.y);
            ks_s[m * HD + d + 1] = bf16r(k1.y * c1.x - k2.y * c1.y);
            ks_s[m * HD + half + d] = bf16r(k2.x * c0.x + k1.x * c0.y);
            ks_s[m * HD + half + d + 1] = bf16r(k2.y * c1.x + k1.y * c1.y);
          }
        }
        dn_cp_wait_all();  // K / V^T fragments (or the staged suffix v rows)
        __syncthreads();
        tick(24);
        tick(25);
        float* po = p.part_o + (long)(h * NCH + c) * 16 * HD;
        float* pml = p.part_ml + (long)(h * NCH + c) * 16 * 2;
        if (prefix) {
          // ---- prefix chunk: keys [key0, key0 + 64) of the cache, tensor cores ----
          const int key0 = c * DN_CK;
          {  // S partial tile: warp (kpart, tslot) -> keys key0 + 8*kpart .. + 8, head-dim groups tslot, tslot + 4
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const bool vlo = g < A, vhi = (g + 8) < A;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int kg = tslot + 4 * u;
              if (kg < HD / 32) {
                const uint4 kf = wme[u * 32];
                const uint4 alo = vlo ? *reinterpret_cast<const uint4*>(q_s + g * ldq + kg * 32 + 8 * t4) : zero4;
                const uint4 ahi = vhi ? *reinterpret_cast<const uint4*>(q_s + (g + 8) * ldq + kg * 32 + 8 * t4) : zero4;
                dn_mma(acc, alo.x, ahi.x, alo.y, ahi.y, kf.x, kf.y);
                dn_mma(acc, alo.z, ahi.z, alo.w, ahi.w, kf.z, kf.w);
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int m = g + (j >> 1) * 8, kk = 8 * kpart + 2 * t4 + (j & 1);
              sp_s[(tslot * 16 + m) * DN_CK + kk] = acc[j];
            }
          }
          __syncthreads();
          tick(26);
          // chunk-local softmax: warp m -> row m; lane -> keys lane, lane+32 (sum of the 4 K-quarter partials, mask)
          if (warp < A) {
            const int m = warp;
            float v0 = 0.f, v1 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              v0 += sp_s[(q * 16 + m) * DN_CK + lane];
              v1 += sp_s[(q * 16 + m) * DN_CK + 32 + lane];
            }
            const int k0 = key0 + lane, k1 = key0 + 32 + lane;
            const bool ok0 = k0 < Pn && ((bits_s[m * 32 + (k0 >> 5)] >> (k0 & 31)) & 1u);
            const bool ok1 = k1 < Pn && ((bits_s[m * 32 + (k1 >> 5)] >> (k1 & 31)) & 1u);
            v0 = ok0 ? v0 : DN_BIG_NEG;
            v1 = ok1 ? v1 : DN_BIG_NEG;
            const float mx = warp_max(fmaxf(v0, v1));
            const float e0 = __expf(v0 - mx), e1 = __expf(v1 - mx);
            const float sum = warp_sum(e0 + e1);
            p_s[m * ldp + lane] = __float2bfloat16_rn(e0);
            p_s[m * ldp + 32 + lane] = __float2bfloat16_rn(e1);
            if (lane == 0) {
              pml[m * 2] = mx;
              pml[m * 2 + 1] = sum;
            }
          }
          __syncthreads();
          tick(27);
          // O_c = P V : warp w -> n8 tile w of the head dims; K = 64 keys (2 groups)
          if (warp < HD / 8) {
            const bool vlo = g < A, vhi = (g + 8) < A;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kg = 0; kg < 2; ++kg) {
              const uint4 vf = wme[(2 + kg) * 32];
              const uint4 alo = vlo ? *reinterpret_cast<const uint4*>(p_s + g * ldp + kg * 32 + 8 * t4) : zero4;
              const uint4 ahi = vhi ? *reinterpret_cast<const uint4*>(p_s + (g + 8) * ldp + kg * 32 + 8 * t4) : zero4;
              dn_mma(acc, alo.x, ahi.x, alo.y, ahi.y, vf.x, vf.y);
              dn_mma(acc, alo.z, ahi.z, alo.w, ahi.w, vf.z, vf.w);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int m = g + (j >> 1) * 8, d = warp * 8 + 2 * t4 + (j & 1);
              if (m < A) po[m * HD + d] = acc[j];
            }
          }
        } else {
          // ---- suffix keys (this step's own A tokens): CUDA cores on shared memory ----
          for (int pr = warp; pr < (r1 - r0) * A; pr += DN_WARPS) {  // logits: warp per (query a, key a2)
            const int a = r0 + pr / A, a2 = pr % A;
            float sacc = 0.f;
            for (int d = lane; d < HD; d += 32) sacc += __bfloat162float(q_s[a * ldq + d]) * ks_s[a2 * HD + d];
            sacc = warp_sum(sacc);
            if (lane == 0) {
              const int key = Pn + a2;
              const bool ok = (bits_s[a * 32 + (key >> 5)] >> (key & 31)) & 1u;
              s_s[a * DN_CK + a2] = ok ? sacc : DN_BIG_NEG;
            }
          }
          __syncthreads();
          if (warp < r1 - r0) {  // warp -> row a; lane a2 -> key a2 (A <= 16)
            const int a = r0 + warp;
            const float v = (lane < A) ? s_s[a * DN_CK + lane] : -3.4e38f;
            const float mx = warp_max(v);
            const float e = (lane < A) ? __expf(v - mx) : 0.f;
            const float sum = warp_sum(e);
            if (lane < A) s_s[a * DN_CK + lane] = bf16r(e);
            if (lane == 0) {
              pml[a * 2] = mx;
              pml[a * 2 + 1] = sum;
            }
          }
          __syncthreads();
          for (int i = threadIdx.x; i < (r1 - r0) * HD; i += DN_THREADS) {
            const int a = r0 + i / HD, d = i % HD;
            float o = 0.f;
            for (int a2 = 0; a2 < A; ++a2) o += s_s[a * DN_CK + a2] * __bfloat162float(vraw[a2 * HD + d]);
            po[a * HD + d] = o;
          }
        }
      }
      tick(4);
      // Wo tile of P3 (one n8 tile per CTA, K split over all 32 warps) -> slots 0,1: the K/V fragments are consumed
      {
        const bf16* Wo = reinterpret_cast<const bf16*>(p.o_w) + (long)l * p.o_ls;
        const bf16* wt3 = o_te > o_tb ? Wo + (long)o_tb * 8 * OD : nullptr;
        bar.arrive();
        dn2_issue_w<2>(wme, wt3, OD, ngo, warp, DN_WARPS);
        bar.wait();
      }
      tick(5);

      // ---------------- P2b: combine the chunks -> O [A, NH*HD]: 4 lanes per 4 outputs, chunks strided over the lanes ----
      {
        const int units = A * (OD >> 2);
        const int per_cta = (units + gridDim.x - 1) / gridDim.x;
        const int quad = threadIdx.x >> 2, sub = threadIdx.x & 3;  // with 32 resident warps issue slots, not latency, are
        for (int u0 = 0; u0 < per_cta; u0 += DN_THREADS / 4) {     // the cost: few threads, 3-5 chunks each
          if (u0 + (warp << 3) >= per_cta) break;                  // (warp-uniform) no live quad in this warp
          const int ui = u0 + quad, unit = blockIdx.x * per_cta + ui;
          const bool valid = ui < per_cta && unit < units;
          const int uu = valid ? unit : 0;
          const int m = uu / (OD >> 2), col = (uu % (OD >> 2)) * 4, h = col / HD, d = col % HD;
          const float* ml = p.part_ml + ((long)h * NCH * 16 + m) * 2;
          const float* oc = p.part_o + ((long)h * NCH * 16 + m) * HD + d;
          float2 mlv[5];
          float4 ov[5];
          float mx = -3.4e38f;
#pragma unroll
          for (int j = 0; j < 5; ++j) {  // chunks sub, sub + 4, ... (NCH <= 17 since Tpad <= 1024)
            const int c = sub + 4 * j;
            mlv[j] = make_float2(-3.4e38f, 0.f);
            ov[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < NCH) {
              mlv[j] = __ldcg(reinterpret_cast<const float2*>(ml + (long)c * 32));
              ov[j] = __ldcg(reinterpret_cast<const float4*>(oc + (long)c * 16 * HD));
            }
          }
#pragma unroll
          for (int j = 0; j < 5; ++j) mx = fmaxf(mx, mlv[j].x);
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          float den = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            if (sub + 4 * j < NCH) {
              const float w = __expf(mlv[j].x - mx);
              den += w * mlv[j].y;
              n0 += w * ov[j].x;
              n1 += w * ov[j].y;
              n2 += w * ov[j].z;
              n3 += w * ov[j].w;
            }
          }
#pragma unroll
          for (int o = 1; o <= 2; o <<= 1) {
            den += __shfl_xor_sync(0xffffffffu, den, o);
            n0 += __shfl_xor_sync(0xffffffffu, n0, o);
            n1 += __shfl_xor_sync(0xffffffffu, n1, o);
            n2 += __shfl_xor_sync(0xffffffffu, n2, o);
            n3 += __shfl_xor_sync(0xffffffffu, n3, o);
          }
          if (valid && sub == 0) {
            uint2 o2;
            o2.x = pack_bf16x2(n0 / den, n1 / den);
            o2.y = pack_bf16x2(n2 / den, n3 / den);
            *reinterpret_cast<uint2*>(Obuf + (long)m * OD + col) = o2;
          }
        }
      }
      tick(21);
      tick(22);
      tick(6);
      bar.sync();
      tick(7);

      // ---------------- P3: XE1 = XE + gate_a * (O Wo^T) ----------------
      const bf16* Wgu = reinterpret_cast<const bf16*>(p.gu_w) + (long)l * p.gu_ls;
      {
        const bf16* Wo = reinterpret_cast<const bf16*>(p.o_w) + (long)l * p.o_ls;
        if (o_te > o_tb) {
          dn_stage(Obuf, OD, h_s, ldo, A, OD);
          dn_cp_wait_all();
          __syncthreads();
          for (int t0 = o_tb; t0 < o_te; ++t0) {
            dn2_mma_warp<2>(h_s, ldo, A, ngo, warp, DN_WARPS, wme, Wo + (long)t0 * 8 * OD, OD, red);
            const bool more = t0 + 1 < o_te;
            if (more) dn2_issue_w<2>(wme, Wo + (long)(t0 + 1) * 8 * OD, OD, ngo, warp, DN_WARPS);
            __syncthreads();
            if (threadIdx.x < 16 * 8) {
              const int m = threadIdx.x >> 3, cc = threadIdx.x & 7;
              if (m < A) {
                const int n = t0 * 8 + cc;
                const float y = bf16r(dn_tile_val(red, 0, m, cc, true));
                const float gt = __bfloat162float(mod_a[2 * D1 + n]);
                const float r = __bfloat162float(xe_s[m * D1 + n]);
                XE1[(long)m * D1 + n] = __float2bfloat16_rn(r + bf16r(y * gt));
              }
            }
            if (more) {
              dn_cp_wait_all();
              __syncthreads();
            }
          }
        }
        tick(8);
        bar.arrive();  // (every warp is done with the staged O: the second weight buffer, behind h, may be filled)
        dn2_issue_w<4>(wme, gu_tile(Wgu, f_pb), D1, ng1, kpart, 8);  // both gate/up passes of P4
        dn2_issue_w<4>(wme2, gu_tile(Wgu, f_pb + 2), D1, ng1, kpart, 8);
        bar.wait();
        tick(9);
      }

      // ---------------- P4: h = adaRMS(XE1); act = gelu(h Wg^T) * (h Wu^T) ----------------
      dn_stage(XE1, D1, xe_s, D1, A, D1);
      dn_cp_wait_all();
      __syncthreads();
      tick(28);
      dn2_ada_norm(xe_s, h_s, ldh, A, D1, mod_f, red);
      __syncthreads();
      tick(29);
      {
        int pass = 0;
        for (int p0 = f_pb; p0 < f_pe; p0 += 2, ++pass) {
          uint4* wb = (pass & 1) ? wme2 : wme;
          dn2_mma_warp<4>(h_s, ldh, A, ng1, kpart, 8, wb, gu_tile(Wgu, p0), D1, red);
          const bool refill = p0 + 4 < f_pe;  // (more than two passes: not at LAP-3B sizes on a full grid)
          if (refill) dn2_issue_w<4>(wb, gu_tile(Wgu, p0 + 4), D1, ng1, kpart, 8);
          __syncthreads();
          if (threadIdx.x < 256) {
            const int e = threadIdx.x;  // 16 rows x 2 pairs x 8 columns = 256 outputs
            const int m = e >> 4, pi = (e >> 3) & 1, cc = e & 7;
            if (m < A && p0 + pi < f_pe) {
              const float gv = bf16r(dn_tile_val(red, 2 * pi, m, cc, false));
              const float uv = bf16r(dn_tile_val(red, 2 * pi + 1, m, cc, false));
              act[(long)m * F1 + (p0 + pi) * 8 + cc] = __float2bfloat16_rn(bf16r(gelu_tanh(gv)) * uv);
            }
          }
          if (p0 + 2 < f_pe) {
            dn_cp_wait_all();  // (nothing pending unless a refill was issued)
            __syncthreads();
          }
        }
      }
      tick(30);
      // P5's tile of Wd (K split over all 32 warps) -> slots 0..3
      {
        const bf16* Wd = reinterpret_cast<const bf16*>(p.down_w) + (long)l * p.down_ls;
        const bf16* wt5 = o_te > o_tb ? Wd + (long)o_tb * 8 * F1 : nullptr;
        bar.arrive();
        dn2_issue_w<4>(wme, wt5, F1, ngf, warp, DN_WARPS);
        tick(10);
        bar.wait();
        tick(11);

        // ---------------- P5: XE = XE1 + gate_f * (act Wd^T) ----------------
        const bool next_w1 = l + 1 < L;  // (the next step's layer 0 is issued in the final phase)
        if (o_te > o_tb) {
          dn_stage(act, F1, h_s, ldf, A, F1);
          dn_cp_wait_all();
          __syncthreads();
          for (int t0 = o_tb; t0 < o_te; ++t0) {
            dn2_mma_warp<4>(h_s, ldf, A, ngf, warp, DN_WARPS, wme, Wd + (long)t0 * 8 * F1, F1, red);
            const bool more = t0 + 1 < o_te;
            if (more) dn2_issue_w<4>(wme, Wd + (long)(t0 + 1) * 8 * F1, F1, ngf, warp, DN_WARPS);
            __syncthreads();
            if (threadIdx.x < 16 * 8) {
              const int m = threadIdx.x >> 3, cc = threadIdx.x & 7;
              if (m < A) {
                const int n = t0 * 8 + cc;
                const float y = bf16r(dn_tile_val(red, 0, m, cc, true));
                const float gt = __bfloat162float(mod_f[2 * D1 + n]);
                const float r = __bfloat162float(xe_s[m * D1 + n]);
                XE[(long)m * D1 + n] = __float2bfloat16_rn(r + bf16r(y * gt));
              }
            }
            if (more) {
              dn_cp_wait_all();
              __syncthreads();
            }
          }
        }
        tick(12);
        bar.arrive();
        if (next_w1) dn2_issue_w<4>(wme, qkv_tile(Wqkv + p.qkv_ls, q_tb), D1, ng1, kpart, 8);
        bar.wait();
        tick(13);
      }
    }

    // ---------------- final: v = action_out_proj(adaRMS(XE)); x += dt * v (lap.py:665-667) ----------------
    dn_stage(XE, D1, xe_s, D1, A, D1);
    dn_stage(mod_s + (long)(p.nm - 1) * 3 * D1, 0, mod_sm, 0, 1, 2 * D1);
    dn_cp_wait_all();
    __syncthreads();
    // the next step's layer-0 Wqkv fragments fly during the final phase and action_in_proj
    if (step + 1 < S) dn2_issue_w<4>(wme, qkv_tile(reinterpret_cast<const bf16*>(p.qkv_w), q_tb), D1, ng1, kpart, 8);
    dn2_ada_norm(xe_s, h_s, ldh, A, D1, mod_sm, red);
    __syncthreads();
    for (int o = warp; o < A * ad; o += DN_WARPS) {
      const int m = o / ad, j = o % ad;
      const float ob = __ldg(p.aout_b + j);  // (in flight during the dot product, not behind the reduction)
      float acc = 0.f;
#pragma unroll 8
      for (int k = lane * 4; k < D1; k += 128) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(p.aout_w + (long)j * D1 + k));
        const uint2 hv = *reinterpret_cast<const uint2*>(h_s + m * ldh + k);
        const float2 h01 = unpack_bf16x2(hv.x), h23 = unpack_bf16x2(hv.y);
        acc += wv.x * h01.x + wv.y * h01.y + wv.z * h23.x + wv.w * h23.y;
      }
      acc = warp_sum(acc);
      if (lane == 0) x_s[o] += p.dt * (acc + ob);
    }
    __syncthreads();
    tick(14);
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < A * ad; i += DN_THREADS) p.x[i] = x_s[i];
  if (PROF && prof_on)
    for (int i = 0; i < 32; ++i) p.prof[(long)blockIdx.x * 32 + i] += prof_s[i];
}

// ======================================================================================================================
// K10c — the same Euler loop as ONE THREAD-BLOCK CLUSTER of 16 CTAs (16 SMs), synchronised by the hardware cluster barrier.
//
// Why: `profiles/r01_denoise_loop.md` — in the 148-CTA kernel above 53 % of every warp's time is spent at the six grid
// barriers per layer (~2 us each plus the imbalance in front of them), while the memory system idles.  A microbenchmark
// (`tools/micro/sm_bw.cu`) shows that 16 SMs alone stream 2.5 TB/s (157 GB/s per SM; 148 SMs share 6.4 TB/s = 43 GB/s each),
// so 16 CTAs are enough to read a layer's 34.6 MB in ~14 us, and `barrier.cluster` costs a few hundred cycles instead of
// ~2 us.  Five phases per layer (the chunk combine is folded into the staging of the o-projection):
//   P1 adaRMS + qkv            : a warp owns whole n8 tiles (full K: no cross-warp reduction, no block barrier)
//   P2 attention               : CTA = (head, half of the prefix keys [+ the suffix keys]); S over up to 512 keys in shared
//                                memory, one softmax over the CTA's keys, P V with one head-dim tile per warp
//   P3 combine halves + o-proj : 8 tiles per CTA x 4 K-quarters (one 16 KB reduction through shared memory)
//   P4 adaRMS + gate/up + GeGLU: a warp owns one (gate, up) tile pair, GeGLU on the accumulator fragments
//   P5 down                    : 8 tiles x 4 K-quarters
// Arithmetic, rounding points and scratch buffers are those of the grid kernel; only the partitioning differs.
//
// STATUS (round 1): correct (same tests as the grid kernel) but NOT yet faster — 27.9 ms per 10 steps against 13.0 ms
// (`LAPB_DENOISE_MODE=cluster` selects it; the default stays the grid kernel).  Per layer-step: P1 16 us, P2 17 us,
// combine 15 us, o-proj 12 us, gate/up 43 us, down 23 us, cluster barriers ~1.2 us each.  The register-fed warp-per-tile
// loops keep only 40-64 KB in flight per SM with 64-byte row segments (20-25 GB/s per SM instead of the 157 GB/s of
// contiguous streaming): the weight slices have to go through a TMA-fed shared-memory ring (contiguous 2-8 KB rows, deep
// prefetch) and the combine needs vector loads before this partitioning can win.
// ======================================================================================================================
constexpr int DNC_CTAS = 16;
constexpr int DNC_MAXK = 512;  // prefix keys per attention CTA (Tpad <= 1024)

__host__ __device__ inline size_t dnc_attn_bytes(int HD) {
  return (size_t)16 * (HD + 8) * 2                 /* q_s   */ + (size_t)16 * (DNC_MAXK + 16) * 4 /* s_s */ +
         (size_t)16 * (DNC_MAXK + 8) * 2           /* p_s   */ + (size_t)16 * HD * 4              /* ks_s */ +
         (size_t)3 * 16 * HD * 2                   /* raw q, k, v */ + (size_t)16 * 16 * 4          /* p of the suffix keys */;
}
__host__ __device__ inline size_t dnc_big_bytes(int D1, int HD, int OD, int F1) {
  size_t b = dn_big_bytes(D1, HD, OD, F1);
  const size_t a = dn_align16(dnc_attn_bytes(HD));
  return b > a ? b : a;
}
__host__ __device__ inline size_t dnc_smem_bytes(int D1, int HD, int OD, int F1) {
  return (size_t)16 * D1 * 2 + dnc_big_bytes(D1, HD, OD, F1) + (size_t)4 * 8 * 32 * 4 * 4 + (size_t)16 * 32 * 4 +
         (size_t)16 * (HD / 2) * 8 + (size_t)6 * D1 * 2 + (size_t)16 * 32 * 4 + 64 + 16;
}

// one n8 weight tile x K groups [kg_begin, kg_end) against the staged rows As: accumulator fragment of this warp
// (lane (g, t): rows g and g + 8, columns 2t and 2t + 1).  The 16-byte fragment of lane (g, t) for K group kg sits at
//   wtile + kg * kstride + g * gstride + 8 t
// row-major weights [N, K]: kstride = 32, gstride = K; TILE-MAJOR packed weights (lapb200_pack_tiles): kstride = 256,
// gstride = 32, i.e. the 32 lanes of a warp read 512 CONTIGUOUS bytes per K group and a whole n8 x K tile is one
// contiguous 16 K-byte range (row-major makes every warp load 8 separate 64-byte row segments, which costs ~4x the
// memory latency: 20-25 GB/s per SM instead of ~150).  U fragments are in flight per lane.
template <int U>
__device__ __forceinline__ void dnc_warp_tile(const bf16* As, int lda, int M, const bf16* wtile, long kstride, long gstride,
                                              int kg_begin, int kg_end, float (&acc)[4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bf16* xlo = As + (long)g * lda + 8 * t;
  const bf16* xhi = As + (long)(g + 8) * lda + 8 * t;
  const bf16* wr = wtile + (long)g * gstride + 8 * t;
  const bool vlo = g < M, vhi = (g + 8) < M;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int kg0 = kg_begin; kg0 < kg_end; kg0 += U) {
    uint4 b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) b[u] = (kg0 + u < kg_end) ? dn_ld_stream(wr + (long)(kg0 + u) * kstride) : zero;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (kg0 + u < kg_end) {
        const uint4 alo = vlo ? *reinterpret_cast<const uint4*>(xlo + (kg0 + u) * 32) : zero;
        const uint4 ahi = vhi ? *reinterpret_cast<const uint4*>(xhi + (kg0 + u) * 32) : zero;
        dn_mma(acc, alo.x, ahi.x, alo.y, ahi.y, b[u].x, b[u].y);
        dn_mma(acc, alo.z, ahi.z, alo.w, ahi.w, b[u].z, b[u].w);
      }
    }
  }
}

__global__ void __launch_bounds__(DN_THREADS, 1) denoise_cluster_kernel(const lapb_denoise_params_t p) {
  extern __shared__ __align__(16) unsigned char dn_smem[];
  const int A = p.A, ad = p.ad, D1 = p.D1, NH = p.NH, HD = p.HD, F1 = p.F1, L = p.L, Pn = p.Pn, W32 = p.W32;
  const int QKV = (NH + 2) * HD, OD = NH * HD, nm3 = p.nm * 3 * D1, S = p.num_steps;
  const int ldh = D1 + 8, ldo = OD + 8, ldf = F1 + 8, ldq = HD + 8, half = HD / 2;
  constexpr int LDS = DNC_MAXK + 16, LDP = DNC_MAXK + 8;
  // ---- shared memory ----
  bf16* xe_s = reinterpret_cast<bf16*>(dn_smem);
  unsigned char* big = dn_smem + (size_t)16 * D1 * 2;
  bf16* h_s = reinterpret_cast<bf16*>(big);
  float* red = reinterpret_cast<float*>(big + dnc_big_bytes(D1, HD, OD, F1));          // [4 K-quarters][8 tiles][32][4]
  float* x_s = red + 4 * 8 * 32 * 4;
  float2* rope_s = reinterpret_cast<float2*>(x_s + 16 * 32);
  bf16* mod_sm = reinterpret_cast<bf16*>(rope_s + 16 * half);
  uint32_t* bits_s = reinterpret_cast<uint32_t*>(mod_sm + 6 * D1);
  float* te_s = reinterpret_cast<float*>(dn_smem);
  // attention scratch (inside `big`)
  bf16* q_s = reinterpret_cast<bf16*>(big);                                            // [16][HD+8]
  float* s_s = reinterpret_cast<float*>(q_s + 16 * ldq);                               // [16][DNC_MAXK+16]
  bf16* p_s = reinterpret_cast<bf16*>(s_s + 16 * LDS);                                 // [16][DNC_MAXK+8]
  float* ks_s = reinterpret_cast<float*>(p_s + 16 * LDP);                              // [16][HD]
  bf16* qraw = reinterpret_cast<bf16*>(ks_s + 16 * HD);
  bf16* kraw = qraw + 16 * HD;
  bf16* vraw = kraw + 16 * HD;
  float* psuf = reinterpret_cast<float*>(vraw + 16 * HD);                              // [16][16] p of the suffix keys

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int rank = blockIdx.x;  // one cluster: rank == cluster_ctarank
  // operand addressing (see dnc_warp_tile): tile-major packed weights / caches, or row-major
  const bool pk = p.packed != 0;
  const long ks = pk ? 256 : 32;
  auto gs = [&](long ld) -> long { return pk ? 32 : ld; };
  const bf16* mod = reinterpret_cast<const bf16*>(p.mod);
  bf16* XE = reinterpret_cast<bf16*>(p.XE);
  bf16* XE1 = reinterpret_cast<bf16*>(p.XE1);
  bf16* qkv = reinterpret_cast<bf16*>(p.qkv);
  bf16* act = reinterpret_cast<bf16*>(p.act);
  const int NCHP = (Pn + DN_CK - 1) / DN_CK;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  const int ng1 = D1 >> 5, ngo = OD >> 5, ngf = F1 >> 5;
  int q_tb, q_te, o_tb, o_te, f_pb, f_pe;
  dn_range(QKV / 8, q_tb, q_te);
  dn_range(D1 / 8, o_tb, o_te);
  dn_range(F1 / 8, f_pb, f_pe);
  unsigned long long prof_last = 0;
  const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  auto tick = [&](int slot) {
    if (prof_on) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (slot >= 0) p.prof[slot] += now - prof_last;
      prof_last = now;
    }
  };
  tick(-1);

  // =========================== prologue: time conditioning of every step ===========================
  {
    const int halfw = D1 / 2;
    for (int i = threadIdx.x; i < S * halfw; i += DN_THREADS) {
      const int r = i / halfw, c = i % halfw;
      const float fraction = (halfw > 1) ? (float)c / (float)(halfw - 1) : 0.f;
      const float period = 4e-3f * powf(4.0f / 4e-3f, fraction);
      const float inp = p.times[r] * (1.0f / period * 2.0f * 3.14159265358979323846f);
      float sn, cs;
      sincosf(inp, &sn, &cs);
      te_s[r * D1 + c] = sn;
      te_s[r * D1 + halfw + c] = cs;
    }
    __syncthreads();
    dn_time_mlp(te_s, S, D1, p.tin_w, p.tin_b, p.s1, nullptr);
    cluster_sync_all();
    for (int i = threadIdx.x; i < S * D1; i += DN_THREADS) te_s[i] = __ldcg(p.s1 + i);
    __syncthreads();
    dn_time_mlp(te_s, S, D1, p.tout_w, p.tout_b, nullptr, reinterpret_cast<bf16*>(p.cond16));
    cluster_sync_all();
    // mod = cond16 @ mod_w^T + mod_b (rows = steps): a warp owns whole tiles
    dn_stage(reinterpret_cast<const bf16*>(p.cond16), D1, h_s, ldh, S, D1);
    dn_cp_wait_all();
    __syncthreads();
    int tb, te;
    dn_range(nm3 / 8, tb, te);
    const bf16* mw = reinterpret_cast<const bf16*>(p.mod_w);
    for (int tile = tb + warp; tile < te; tile += DN_WARPS) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      dnc_warp_tile<8>(h_s, ldh, S, mw + (long)tile * 8 * D1, 32, D1, 0, ng1, acc);  // mod_w stays row-major
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = g + (j >> 1) * 8, n = tile * 8 + 2 * t4 + (j & 1);
        if (m < S) reinterpret_cast<bf16*>(p.mod)[(long)m * nm3 + n] = __float2bfloat16_rn(bf16r(acc[j]) + bf16r(p.mod_b[n]));
      }
    }
    for (int i = threadIdx.x; i < A * ad; i += DN_THREADS) x_s[i] = p.x[i];
    for (int i = threadIdx.x; i < A * half; i += DN_THREADS) {
      const int m = i / half, d = i % half;
      float sn, cs;
      sincosf((float)p.pos[m] / p.timescale[d], &sn, &cs);
      rope_s[i] = make_float2(cs, sn);
    }
    for (int i = threadIdx.x; i < A * W32; i += DN_THREADS) bits_s[(i / W32) * 32 + (i % W32)] = p.bits[i];
    cluster_sync_all();
    tick(0);
  }

  // attention role of this CTA: (head, half of the prefix chunks); the second half also owns the suffix keys
  const int at_h = rank >> 1, at_half = rank & 1;
  const int c_first = (NCHP + 1) / 2;
  const int c_begin = at_half ? c_first : 0, c_end = at_half ? NCHP : c_first;
  const int key0 = c_begin * DN_CK, nk = (c_end - c_begin) * DN_CK;  // nk <= DNC_MAXK
  const bool at_active = at_h < NH;

  for (int step = 0; step < S; ++step) {
    const bf16* mod_s = mod + (long)step * nm3;
    for (int i = threadIdx.x; i < A * D1; i += DN_THREADS) {
      const int m = i / D1, n = i % D1;
      float v = 0.f;
      for (int j = 0; j < ad; ++j) v += x_s[m * ad + j] * __ldg(p.ain_w + (long)n * ad + j);
      xe_s[i] = __float2bfloat16_rn(v + __ldg(p.ain_b + n));
    }
    __syncthreads();
    tick(1);

    for (int l = 0; l < L; ++l) {
      const bf16* Wqkv = reinterpret_cast<const bf16*>(p.qkv_w) + (long)l * p.qkv_ls;
      const bf16* Wo = reinterpret_cast<const bf16*>(p.o_w) + (long)l * p.o_ls;
      const bf16* Wgu = reinterpret_cast<const bf16*>(p.gu_w) + (long)l * p.gu_ls;
      const bf16* Wd = reinterpret_cast<const bf16*>(p.down_w) + (long)l * p.down_ls;
      const bf16* Kc = reinterpret_cast<const bf16*>(p.Kc) + (long)l * p.kc_ls;
      const bf16* VcT = reinterpret_cast<const bf16*>(p.VcT) + (long)l * p.vct_ls;
      const bf16* mod_a = mod_sm;
      const bf16* mod_f = mod_sm + 3 * D1;

      // ---------------- P1: h = adaRMS(XE); qkv = h Wqkv^T (warp = tile) ----------------
      if (l > 0) dn_stage(XE, D1, xe_s, D1, A, D1);
      dn_stage(mod_s + (long)(2 * l) * 3 * D1, 0, mod_sm, 0, 1, 6 * D1);
      dn_cp_wait_all();
      __syncthreads();
      dn_ada_norm(xe_s, h_s, ldh, A, D1, mod_a);
      __syncthreads();
      for (int tile = q_tb + warp; tile < q_te; tile += DN_WARPS) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        dnc_warp_tile<8>(h_s, ldh, A, Wqkv + (long)tile * 8 * D1, ks, gs(D1), 0, ng1, acc);
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const int m = g + (j >> 1) * 8;
          if (m < A)
            *reinterpret_cast<uint32_t*>(qkv + (long)m * QKV + tile * 8 + 2 * t4) = pack_bf16x2(acc[j], acc[j + 1]);
        }
      }
      tick(2);
      cluster_sync_all();
      tick(3);

      // ---------------- P2: attention partial of (head, key half) ----------------
      if (at_active) {
        dn_stage(qkv + at_h * HD, QKV, qraw, HD, A, HD);
        if (at_half) {
          dn_stage(qkv + NH * HD, QKV, kraw, HD, A, HD);
          dn_stage(qkv + (NH + 1) * HD, QKV, vraw, HD, A, HD);
        }
        dn_cp_wait_all();
        __syncthreads();
        for (int i = threadIdx.x; i < A * half; i += DN_THREADS) {
          const int m = i / half, d = i % half;
          const float cs = rope_s[i].x, sn = rope_s[i].y;
          const float x1 = __bfloat162float(qraw[m * HD + d]), x2 = __bfloat162float(qraw[m * HD + half + d]);
          q_s[m * ldq + d] = __float2bfloat16_rn(bf16r(x1 * cs - x2 * sn) * p.qscale);
          q_s[m * ldq + half + d] = __float2bfloat16_rn(bf16r(x2 * cs + x1 * sn) * p.qscale);
          if (at_half) {
            const float k1 = __bfloat162float(kraw[m * HD + d]), k2 = __bfloat162float(kraw[m * HD + half + d]);
            ks_s[m * HD + d] = bf16r(k1 * cs - k2 * sn);
            ks_s[m * HD + half + d] = bf16r(k2 * cs + k1 * sn);
          }
        }
        __syncthreads();
        // S = Q_h K^T over this CTA's prefix keys: a warp owns 8-key tiles (full K = HD)
        for (int tl = warp; tl < nk / 8; tl += DN_WARPS) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          dnc_warp_tile<8>(q_s, ldq, A, Kc + (long)(key0 + 8 * tl) * HD, ks, gs(HD), 0, HD / 32, acc);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int m = g + (j >> 1) * 8, kk = 8 * tl + 2 * t4 + (j & 1), key = key0 + kk;
            bool ok = false;
            if (m < A && key < Pn) ok = (bits_s[m * 32 + (key >> 5)] >> (key & 31)) & 1u;
            s_s[m * LDS + kk] = ok ? acc[j] : DN_BIG_NEG;
          }
        }
        if (at_half) {  // logits of the suffix keys: warp per (query a, key a2)
          for (int pr = warp; pr < A * A; pr += DN_WARPS) {
            const int a = pr / A, a2 = pr % A;
            float sacc = 0.f;
            for (int d = lane; d < HD; d += 32) sacc += __bfloat162float(q_s[a * ldq + d]) * ks_s[a2 * HD + d];
            sacc = warp_sum(sacc);
            if (lane == 0) {
              const int key = Pn + a2;
              const bool ok = (bits_s[a * 32 + (key >> 5)] >> (key & 31)) & 1u;
              s_s[a * LDS + nk + a2] = ok ? sacc : DN_BIG_NEG;
            }
          }
        }
        __syncthreads();
        // softmax over the CTA's keys: warp m -> row m
        float* pml = p.part_ml + (long)rank * 16 * 2;
        if (warp < A) {
          const int m = warp, ncol = nk + (at_half ? A : 0);
          float mx = -3.4e38f;
          for (int c = lane; c < ncol; c += 32) mx = fmaxf(mx, s_s[m * LDS + c]);
          mx = warp_max(mx);
          float sum = 0.f;
          for (int c = lane; c < ncol; c += 32) {
            const float e = __expf(s_s[m * LDS + c] - mx);
            sum += e;
            if (c < nk) p_s[m * LDP + c] = __float2bfloat16_rn(e);
            else psuf[m * 16 + (c - nk)] = bf16r(e);
          }
          sum = warp_sum(sum);
          if (lane == 0) {
            pml[m * 2] = mx;
            pml[m * 2 + 1] = sum;
          }
        }
        __syncthreads();
        // O_partial = P V: warp w -> head-dim tile w; prefix keys by MMA from V^T, suffix keys on the CUDA cores
        float* po = p.part_o + (long)rank * 16 * HD;
        if (warp < HD / 8) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          dnc_warp_tile<8>(p_s, LDP, A, VcT + (long)(warp * 8) * p.TpadK + (pk ? (long)(key0 / 32) * 256 : (long)key0), ks,
                           gs(p.TpadK), 0, nk / 32, acc);
          if (at_half) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int m = g + (j >> 1) * 8, d = warp * 8 + 2 * t4 + (j & 1);
              if (m < A) {
                float o = 0.f;
                for (int a2 = 0; a2 < A; ++a2) o += psuf[m * 16 + a2] * __bfloat162float(vraw[a2 * HD + d]);
                acc[j] += o;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int m = g + (j >> 1) * 8, d = warp * 8 + 2 * t4 + (j & 1);
            if (m < A) po[m * HD + d] = acc[j];
          }
        }
      }
      tick(4);
      cluster_sync_all();
      tick(5);

      // ---------------- P3: combine the two halves of every head -> O (staged), XE1 = XE + gate_a * (O Wo^T) ----------------
      for (int i = threadIdx.x; i < A * (OD / 4); i += DN_THREADS) {
        const int m = i / (OD / 4), c = (i - m * (OD / 4)) * 4, h = c / HD, d = c - h * HD;
        const float2 ml0 = __ldcg(reinterpret_cast<const float2*>(p.part_ml + ((long)(2 * h) * 16 + m) * 2));
        const float2 ml1 = __ldcg(reinterpret_cast<const float2*>(p.part_ml + ((long)(2 * h + 1) * 16 + m) * 2));
        const float4 o0 = __ldcg(reinterpret_cast<const float4*>(p.part_o + ((long)(2 * h) * 16 + m) * HD + d));
        const float4 o1 = __ldcg(reinterpret_cast<const float4*>(p.part_o + ((long)(2 * h + 1) * 16 + m) * HD + d));
        const float mx = fmaxf(ml0.x, ml1.x);
        const float w0 = __expf(ml0.x - mx), w1 = __expf(ml1.x - mx);
        const float inv = 1.0f / (w0 * ml0.y + w1 * ml1.y);
        uint2 u;
        u.x = pack_bf16x2((w0 * o0.x + w1 * o1.x) * inv, (w0 * o0.y + w1 * o1.y) * inv);
        u.y = pack_bf16x2((w0 * o0.z + w1 * o1.z) * inv, (w0 * o0.w + w1 * o1.w) * inv);
        *reinterpret_cast<uint2*>(h_s + m * ldo + c) = u;
      }
      __syncthreads();
      tick(6);
      {
        const int tl = warp & 7, kq = warp >> 3;  // 8 tiles x 4 K-quarters
        for (int t0 = o_tb; t0 < o_te; t0 += 8) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          const int gq = (ngo + 3) / 4;
          if (t0 + tl < o_te)
            dnc_warp_tile<8>(h_s, ldo, A, Wo + (long)(t0 + tl) * 8 * OD, ks, gs(OD), kq * gq, min(ngo, (kq + 1) * gq), acc);
          *reinterpret_cast<float4*>(red + (((kq * 8 + tl) * 32 + lane) << 2)) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          __syncthreads();
          {
            const int e = threadIdx.x, et = e >> 7, m = (e >> 3) & 15, cc = e & 7;
            if (m < A && t0 + et < o_te) {
              const int src_lane = (m & 7) * 4 + (cc >> 1), idx = (m >> 3) * 2 + (cc & 1);
              float v = 0.f;
#pragma unroll
              for (int q = 0; q < 4; ++q) v += red[((q * 8 + et) * 32 + src_lane) * 4 + idx];
              const int n = (t0 + et) * 8 + cc;
              const float y = bf16r(v);
              const float gt = __bfloat162float(mod_a[2 * D1 + n]);
              const float r = __bfloat162float(xe_s[m * D1 + n]);
              XE1[(long)m * D1 + n] = __float2bfloat16_rn(r + bf16r(y * gt));
            }
          }
          __syncthreads();
        }
      }
      tick(8);
      cluster_sync_all();
      tick(9);

      // ---------------- P4: h = adaRMS(XE1); act = gelu(h Wg^T) * (h Wu^T) (warp = gate/up tile pair) ----------------
      dn_stage(XE1, D1, xe_s, D1, A, D1);
      dn_cp_wait_all();
      __syncthreads();
      dn_ada_norm(xe_s, h_s, ldh, A, D1, mod_f);
      __syncthreads();
      for (int pr = f_pb + warp; pr < f_pe; pr += DN_WARPS) {
        float ag[4] = {0.f, 0.f, 0.f, 0.f}, au[4] = {0.f, 0.f, 0.f, 0.f};
        dnc_warp_tile<8>(h_s, ldh, A, Wgu + (long)pr * 8 * D1, ks, gs(D1), 0, ng1, ag);
        dnc_warp_tile<8>(h_s, ldh, A, Wgu + ((long)F1 + (long)pr * 8) * D1, ks, gs(D1), 0, ng1, au);
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const int m = g + (j >> 1) * 8;
          if (m < A) {
            const float a0 = bf16r(gelu_tanh(bf16r(ag[j]))) * bf16r(au[j]);
            const float a1 = bf16r(gelu_tanh(bf16r(ag[j + 1]))) * bf16r(au[j + 1]);
            *reinterpret_cast<uint32_t*>(act + (long)m * F1 + pr * 8 + 2 * t4) = pack_bf16x2(a0, a1);
          }
        }
      }
      tick(10);
      cluster_sync_all();
      tick(11);

      // ---------------- P5: XE = XE1 + gate_f * (act Wd^T) ----------------
      dn_stage(act, F1, h_s, ldf, A, F1);
      dn_cp_wait_all();
      __syncthreads();
      {
        const int tl = warp & 7, kq = warp >> 3;
        for (int t0 = o_tb; t0 < o_te; t0 += 8) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          const int gq = (ngf + 3) / 4;
          if (t0 + tl < o_te)
            dnc_warp_tile<8>(h_s, ldf, A, Wd + (long)(t0 + tl) * 8 * F1, ks, gs(F1), kq * gq, min(ngf, (kq + 1) * gq), acc);
          *reinterpret_cast<float4*>(red + (((kq * 8 + tl) * 32 + lane) << 2)) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          __syncthreads();
          {
            const int e = threadIdx.x, et = e >> 7, m = (e >> 3) & 15, cc = e & 7;
            if (m < A && t0 + et < o_te) {
              const int src_lane = (m & 7) * 4 + (cc >> 1), idx = (m >> 3) * 2 + (cc & 1);
              float v = 0.f;
#pragma unroll
              for (int q = 0; q < 4; ++q) v += red[((q * 8 + et) * 32 + src_lane) * 4 + idx];
              const int n = (t0 + et) * 8 + cc;
              const float y = bf16r(v);
              const float gt = __bfloat162float(mod_f[2 * D1 + n]);
              const float r = __bfloat162float(xe_s[m * D1 + n]);
              XE[(long)m * D1 + n] = __float2bfloat16_rn(r + bf16r(y * gt));
            }
          }
          __syncthreads();
        }
      }
      tick(12);
      cluster_sync_all();
      tick(13);
    }

    // ---------------- final: v = action_out_proj(adaRMS(XE)); x += dt * v ----------------
    dn_stage(XE, D1, xe_s, D1, A, D1);
    dn_stage(mod_s + (long)(p.nm - 1) * 3 * D1, 0, mod_sm, 0, 1, 2 * D1);
    dn_cp_wait_all();
    __syncthreads();
    dn_ada_norm(xe_s, h_s, ldh, A, D1, mod_sm);
    __syncthreads();
    for (int o = warp; o < A * ad; o += DN_WARPS) {
      const int m = o / ad, j = o % ad;
      float acc = 0.f;
#pragma unroll 8
      for (int k = lane; k < D1; k += 32) acc += __ldg(p.aout_w + (long)j * D1 + k) * __bfloat162float(h_s[m * ldh + k]);
      acc = warp_sum(acc);
      if (lane == 0) x_s[o] += p.dt * (acc + __ldg(p.aout_b + j));
    }
    __syncthreads();
    tick(14);
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < A * ad; i += DN_THREADS) p.x[i] = x_s[i];
}

// V^T of the prefix part of the KV cache: VcT[l][d][j] = Vc[l][j][d], j < TpadK (zero beyond Pn) — written once per
// inference after the prefix pass so that P V in the denoise loop reads keys contiguously (mma.sync B operand).
__global__ void transpose_v_kernel(const bf16* __restrict__ Vc, bf16* __restrict__ VcT, int Tpad, int TpadK, int HD, int Pn) {
  __shared__ bf16 tile[32][33];
  const int l = blockIdx.z;
  const int j0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const bf16* src = Vc + (long)l * Tpad * HD;
  bf16* dst = VcT + (long)l * HD * TpadK;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = j0 + r, d = d0 + threadIdx.x;
    tile[r][threadIdx.x] = (j < Pn && j < Tpad && d < HD) ? src[(long)j * HD + d] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, j = j0 + threadIdx.x;
    if (d < HD && j < TpadK) dst[(long)d * TpadK + j] = tile[threadIdx.x][r];
  }
}

// Tile-major packing for the weight-streaming kernels: element (r, c) of a [rows, cols] matrix (rows % 8 == 0,
// cols % 32 == 0) moves to  ((r/8 * cols/32 + c/32) * 8 + r%8) * 32 + c%32 : every n8 x K tile is one contiguous
// range and the 32 lanes of a warp (lane = 4 (r%8) + t, 8 elements each) read 512 contiguous bytes per K group.
// Source element (r, c) is src[r * rs + c * cs] (cs != 1: a transposed source, e.g. V^T from the value cache), zero
// for c >= valid_cols.  grid.y = batch (layers).
__global__ void pack_tiles_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long rows, long cols, long rs, long cs,
                                  long valid_cols, long src_bs, long dst_bs) {
  const bf16* sb = src + (long)blockIdx.y * src_bs;
  bf16* db = dst + (long)blockIdx.y * dst_bs;
  const long n = rows * cols;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long)gridDim.x * blockDim.x) {
    const long e = o & 31, g = (o >> 5) & 7, blk = o >> 8;
    const long ng = cols >> 5, kg = blk % ng, tile = blk / ng;
    const long r = tile * 8 + g, c = kg * 32 + e;
    db[o] = (c < valid_cols) ? sb[r * rs + c * cs] : __float2bfloat16_rn(0.f);
  }
}

}  // namespace lapb

using namespace lapb;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int lapb200_denoise_supported(int64_t B, int64_t A, int64_t ad, int64_t D1, int64_t NH, int64_t HD, int64_t F1,
                              int64_t Pn, int64_t Tpad, int64_t num_steps) {
  return B == 1 && A >= 1 && A <= 16 && ad >= 1 && ad <= 32 && num_steps >= 1 && num_steps <= 16 && D1 % 32 == 0 &&
         D1 >= 64 && D1 <= 2048 && HD % 32 == 0 && HD >= 32 && HD <= 256 && F1 % 32 == 0 && (NH * HD) % 32 == 0 &&
         NH >= 1 && Pn >= 1 && Tpad >= ((Pn + DN_CK - 1) / DN_CK) * DN_CK && Tpad % 32 == 0 && Tpad <= 1024 &&
         (dn_smem_bytes((int)D1, (int)HD, (int)(NH * HD), (int)F1) <= (size_t)227 * 1024 ||
          dn2_smem_bytes((int)A, (int)num_steps, (int)D1, (int)HD, (int)(NH * HD), (int)F1) <= (size_t)227 * 1024);
}

int lapb200_denoise_grid(void) { return num_sms(); }

int lapb200_transpose_v(const void* Vc, void* VcT, int64_t L, int64_t Tpad, int64_t TpadK, int64_t HD, int64_t Pn,
                        lapb_stream_t s) {
  LAPB_REQUIRE(TpadK % 8 == 0 && TpadK <= Tpad, "transpose_v: TpadK must be a multiple of 8 and <= Tpad");
  dim3 grid(cdiv(TpadK, 32), cdiv(HD, 32), (unsigned)L), block(32, 8);
  transpose_v_kernel<<<grid, block, 0, STREAM(s)>>>((const bf16*)Vc, (bf16*)VcT, (int)Tpad, (int)TpadK, (int)HD, (int)Pn);
  LAPB_LAUNCH_OK("transpose_v");
  return 0;
}

int lapb200_pack_tiles(const void* src, void* dst, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride,
                       int64_t valid_cols, int64_t batch, int64_t src_bs, int64_t dst_bs, lapb_stream_t s) {
  LAPB_REQUIRE(rows % 8 == 0 && cols % 32 == 0 && batch >= 1, "pack_tiles: rows %% 8, cols %% 32 (got %ld x %ld)", (long)rows,
               (long)cols);
  const long n = rows * cols;
  dim3 grid((unsigned)std::min<long>((n + 255) / 256, 4096), (unsigned)batch);
  pack_tiles_kernel<<<grid, 256, 0, STREAM(s)>>>((const bf16*)src, (bf16*)dst, rows, cols, row_stride, col_stride, valid_cols,
                                                 src_bs, dst_bs);
  LAPB_LAUNCH_OK("pack_tiles");
  return 0;
}

int lapb200_denoise_loop(const lapb_denoise_params_t* params, lapb_stream_t s) {
  const lapb_denoise_params_t& p = *params;
  LAPB_REQUIRE(lapb200_denoise_supported(1, p.A, p.ad, p.D1, p.NH, p.HD, p.F1, p.Pn, p.Tpad, p.num_steps),
               "denoise_loop: unsupported shape (A=%d ad=%d D1=%d NH=%d HD=%d F1=%d Pn=%d Tpad=%d steps=%d)", p.A, p.ad,
               p.D1, p.NH, p.HD, p.F1, p.Pn, p.Tpad, p.num_steps);
  LAPB_REQUIRE(p.TpadK == ((p.Pn + DN_CK - 1) / DN_CK) * DN_CK, "denoise_loop: TpadK must be round_up(Pn, 64)");
  // Two partitionings of the same loop: "cluster" = one 16-CTA thread-block cluster (hardware cluster barriers; needs
  // 8 query heads so that (head, key half) maps onto the 16 CTAs), "grid" = one CTA per SM with grid barriers.
  static int mode = -1;  // 0 grid v2 (v1 when its shared memory does not fit), 1 cluster, 2 grid v1
  if (mode < 0) {
    const char* e = getenv("LAPB_DENOISE_MODE");
    mode = (e && e[0] == 'c') ? 1 : (e && strcmp(e, "grid1") == 0) ? 2 : 0;  // default: grid (measured faster than K10c)
  }
  LAPB_REQUIRE(!p.packed || mode == 1, "denoise_loop: tile-major packed operands are read by the cluster kernel only "
               "(LAPB_DENOISE_MODE=cluster)");
  const size_t csmem = dnc_smem_bytes(p.D1, p.HD, p.NH * p.HD, p.F1);
  if (mode == 1 && p.NH * 2 <= DNC_CTAS && p.TpadK <= 2 * DNC_MAXK && csmem <= (size_t)227 * 1024) {
    static size_t cconf = 0;
    if (csmem > cconf) {
      LAPB_CUDA_OK(cudaFuncSetAttribute(denoise_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      LAPB_CUDA_OK(cudaFuncSetAttribute(denoise_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      cconf = csmem;
    }
    cudaLaunchConfig_t ccfg = {};
    ccfg.gridDim = dim3(DNC_CTAS);
    ccfg.blockDim = dim3(DN_THREADS);
    ccfg.dynamicSmemBytes = csmem;
    ccfg.stream = STREAM(s);
    cudaLaunchAttribute cattr[1];
    cattr[0].id = cudaLaunchAttributeClusterDimension;
    cattr[0].val.clusterDim.x = DNC_CTAS;
    cattr[0].val.clusterDim.y = 1;
    cattr[0].val.clusterDim.z = 1;
    ccfg.attrs = cattr;
    ccfg.numAttrs = 1;
    const cudaError_t ce = cudaLaunchKernelEx(&ccfg, denoise_cluster_kernel, p);
    if (ce == cudaSuccess) return 0;
    (void)cudaGetLastError();  // e.g. no GPC with 16 free SMs for the cluster: use the grid kernel from now on
    LAPB_REQUIRE(!p.packed, "denoise_loop: the cluster launch failed (%s) and the grid kernel cannot read packed operands",
                 cudaGetErrorString(ce));
    mode = 0;
  }
  const size_t smem2 = dn2_smem_bytes(p.A, p.num_steps, p.D1, p.HD, p.NH * p.HD, p.F1);
  const bool v2 = mode != 2 && !(p.flags & 128) && smem2 <= (size_t)227 * 1024;  // flags bit 7: round-1 layout (tests)
  const size_t smem = v2 ? smem2 : dn_smem_bytes(p.D1, p.HD, p.NH * p.HD, p.F1);
  LAPB_REQUIRE(smem <= 227 * 1024, "denoise_loop: needs %zu bytes of shared memory (> 227 KB)", smem);
  const void* kern = v2 ? (p.prof ? (const void*)denoise_loop2_kernel<true> : (const void*)denoise_loop2_kernel<false>)
                        : (const void*)denoise_loop_kernel;
  static size_t configured[3] = {0, 0, 0};
  const int ki = v2 ? (p.prof ? 2 : 1) : 0;
  if (smem > configured[ki]) {
    LAPB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[ki] = smem;
  }
  int per_sm = 0;
  LAPB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DN_THREADS, smem));
  LAPB_REQUIRE(per_sm >= 1, "denoise_loop: kernel does not fit on an SM (smem %zu)", smem);
  LAPB_CUDA_OK(cudaMemsetAsync(p.sync, 0, LAPB_DENOISE_SYNC_WORDS * sizeof(uint32_t), STREAM(s)));
  cudaLaunchConfig_t cfg = {};
  // Grid: one CTA per SM by default.  128 CTAs balance LAP-3B's phases exactly (128 o/down tiles, 512 gate/up pairs);
  // LAPB_DENOISE_CTAS overrides for experiments.
  int ctas = num_sms();
  if (const char* e = getenv("LAPB_DENOISE_CTAS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= num_sms()) ctas = v;
  }
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(DN_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = STREAM(s);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int fold_env = -1;
  if (fold_env < 0) {
    const char* e = getenv("LAPB_DENOISE_FOLD");
    fold_env = (e && atoi(e) != 0 && p.NH <= LAPB_DENOISE_SYNC_WORDS - 2) ? 1 : 0;
  }
  if (v2) {
    void* args[] = {(void*)&p};
    LAPB_CUDA_OK(cudaLaunchKernelExC(&cfg, kern, args));
  } else {
    LAPB_CUDA_OK(cudaLaunchKernelEx(&cfg, denoise_loop_kernel, p, fold_env));
  }
  return 0;
}

}  // extern "C"
