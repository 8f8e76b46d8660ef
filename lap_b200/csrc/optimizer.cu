// K12 — fused optimizer over the flat fp32 state (scripts/train.py:363-415, OP/training/optimizer.py:76-85):
//   global-norm clip -> AdamW -> EMA -> bf16 compute copy, plus grad / param norms, in two launches:
//   (1) sumsq_partials: per-CTA partial sums of g^2                      (reads g once)
//   (2) adamw_ema: each CTA re-reduces the partials (tiny), derives the clip scale on device (no host sync),
//       then streams p,g,m,v,ema once: 5 fp32 reads + 4 fp32 writes + 1 bf16 write per element.  HBM-bound.
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"

namespace lapb {

typedef __nv_bfloat16 bf16;
constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS)
sumsq_partials_kernel(const float* __restrict__ x, long n, float* __restrict__ partials) {
  __shared__ float red[32];
  float acc = 0.f;
  long i = ((long)blockIdx.x * OPT_THREADS + threadIdx.x) * 4;
  long stride = (long)gridDim.x * OPT_THREADS * 4;
  for (; i + 4 <= n; i += stride) {
    float4 v = *reinterpret_cast<const float4*>(x + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long t = n & ~3L; t < n; ++t) acc += x[t] * x[t];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

struct AdamArgs {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* ema;        // may be null
  bf16* w16;         // may be null
  long n;
  const float* gpartials;  // [n_partials] partial sums of g^2
  int n_partials;
  float* stats;      // [0] grad_norm  [1] clip scale  [2] += sum p_new^2 over the kernel range
  long kernel_begin, kernel_end;  // element range counted in param_norm
  float lr, b1, b2, eps, wd, bc1, bc2, clip, ema_decay;
  int ema_on;
  const float* hyper;  // optional device array [lr, bc1, bc2, ema_decay, ema_on]: overrides the by-value fields
};

__global__ void __launch_bounds__(OPT_THREADS) adamw_ema_kernel(AdamArgs a) {
  __shared__ float red[32];
  if (a.hyper) {  // per-step scalars live in device memory so a captured CUDA graph can be replayed unchanged
    a.lr = a.hyper[0]; a.bc1 = a.hyper[1]; a.bc2 = a.hyper[2]; a.ema_decay = a.hyper[3];
    a.ema_on = a.hyper[4] != 0.f;
  }
  __shared__ float s_scale;
  float acc = 0.f;
  for (int i = threadIdx.x; i < a.n_partials; i += OPT_THREADS) acc += a.gpartials[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    float gnorm = sqrtf(acc);
    // optax.clip_by_global_norm: g if norm < max_norm else g / norm * max_norm
    float sc = (gnorm < a.clip) ? 1.0f : a.clip / gnorm;
    s_scale = sc;
    if (blockIdx.x == 0) {
      a.stats[0] = gnorm;
      a.stats[1] = sc;
    }
  }
  __syncthreads();
  const float scale = s_scale;
  float pn = 0.f;
  long i = ((long)blockIdx.x * OPT_THREADS + threadIdx.x) * 4;
  long stride = (long)gridDim.x * OPT_THREADS * 4;
  for (; i + 4 <= a.n; i += stride) {
    float4 p4 = *reinterpret_cast<const float4*>(a.p + i);
    float4 g4 = *reinterpret_cast<const float4*>(a.g + i);
    float4 m4 = *reinterpret_cast<const float4*>(a.m + i);
    float4 v4 = *reinterpret_cast<const float4*>(a.v + i);
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float g = gg[j] * scale;
      mm[j] = a.b1 * mm[j] + (1.0f - a.b1) * g;
      vv[j] = a.b2 * vv[j] + (1.0f - a.b2) * g * g;
      float upd = (mm[j] / a.bc1) / (sqrtf(vv[j] / a.bc2) + a.eps) + a.wd * pp[j];
      pp[j] = pp[j] - a.lr * upd;
    }
    *reinterpret_cast<float4*>(a.p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(a.m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(a.v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (a.ema && a.ema_on) {
      float4 e4 = *reinterpret_cast<const float4*>(a.ema + i);
      e4.x = a.ema_decay * e4.x + (1.0f - a.ema_decay) * pp[0];
      e4.y = a.ema_decay * e4.y + (1.0f - a.ema_decay) * pp[1];
      e4.z = a.ema_decay * e4.z + (1.0f - a.ema_decay) * pp[2];
      e4.w = a.ema_decay * e4.w + (1.0f - a.ema_decay) * pp[3];
      *reinterpret_cast<float4*>(a.ema + i) = e4;
    }
    if (a.w16) {
      uint2 o;
      o.x = pack_bf16x2(pp[0], pp[1]);
      o.y = pack_bf16x2(pp[2], pp[3]);
      *reinterpret_cast<uint2*>(a.w16 + i) = o;
    }
    if (i >= a.kernel_begin && i < a.kernel_end) pn += pp[0] * pp[0] + pp[1] * pp[1] + pp[2] * pp[2] + pp[3] * pp[3];
  }
  pn = block_sum(pn, red);
  if (threadIdx.x == 0 && pn != 0.f) atomicAdd(a.stats + 2, pn);
}

}  // namespace lapb

using namespace lapb;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int lapb200_opt_num_partials(void) { return num_sms() * 8; }

int lapb200_sumsq_partials(const float* x, int64_t n, float* partials, lapb_stream_t s) {
  sumsq_partials_kernel<<<num_sms() * 8, OPT_THREADS, 0, STREAM(s)>>>(x, n, partials);
  LAPB_LAUNCH_OK("sumsq_partials");
  return 0;
}

// n must be a multiple of 4 (the flat state is padded). stats: float[4], stats[2] must be zeroed by the caller.
int lapb200_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, void* w16, int64_t n,
                      const float* gpartials, int64_t n_partials, float* stats, int64_t kernel_begin,
                      int64_t kernel_end, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2,
                      float clip, float ema_decay, int64_t ema_on, const float* hyper, lapb_stream_t s) {
  LAPB_REQUIRE(n % 4 == 0, "adamw: n must be a multiple of 4");
  LAPB_REQUIRE(kernel_begin % 4 == 0 && kernel_end % 4 == 0, "adamw: kernel range must be 4-aligned");
  AdamArgs a;
  a.p = p; a.g = g; a.m = m; a.v = v; a.ema = ema; a.w16 = (bf16*)w16; a.n = n;
  a.gpartials = gpartials; a.n_partials = (int)n_partials; a.stats = stats;
  a.kernel_begin = kernel_begin; a.kernel_end = kernel_end;
  a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps; a.wd = wd; a.bc1 = bc1; a.bc2 = bc2; a.clip = clip;
  a.ema_decay = ema_decay; a.ema_on = (int)ema_on; a.hyper = hyper;
  adamw_ema_kernel<<<num_sms() * 8, OPT_THREADS, 0, STREAM(s)>>>(a);
  LAPB_LAUNCH_OK("adamw_ema");
  return 0;
}

}  // extern "C"
