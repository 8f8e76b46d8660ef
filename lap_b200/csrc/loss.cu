// K9 (loss half) + a20: language cross-entropy over fp32 LM-head logits and the flow-matching MSE.
//   lap.py:221-260  logp = log_softmax(logits); tok = logp[target]; per-sample masked mean; weighted sum
//   lap.py:291-301  v = action_out_proj(suffix_out); mean((v-u)^2) over (A, ad)
// The CE kernel reads each logits row twice (online max/sum, then gradient) and writes bf16 dlogits in the same
// launch: rows are ~1 MB of fp32 so the second read is served from L2.  HBM-bound.
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"

namespace lapb {

typedef __nv_bfloat16 bf16;

// one CTA (1024 threads) per row.
//   nll[r] = logsumexp(row) - row[target];  dlogits[r, j] = w[r] * (softmax_j - [j == target])  (bf16)
__global__ void __launch_bounds__(1024)
ce_fwd_bwd_kernel(const float* __restrict__ logits, long ld, const int* __restrict__ targets,
                  const float* __restrict__ weights, float* __restrict__ nll, bf16* __restrict__ dlogits, long ldd,
                  int V) {
  __shared__ float red[32];
  long r = blockIdx.x;
  const float* row = logits + r * ld;
  float w = weights[r];
  // online logsumexp
  float m = -3.4e38f, s = 0.f;
  for (int c = threadIdx.x * 4; c < V; c += 1024 * 4) {
    float4 x = *reinterpret_cast<const float4*>(row + c);
    float lm = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
    if (lm > m) {
      s *= __expf(m - lm);
      m = lm;
    }
    s += __expf(x.x - m) + __expf(x.y - m) + __expf(x.z - m) + __expf(x.w - m);
  }
  float gm = block_max(m, red);
  s *= __expf(m - gm);
  s = block_sum(s, red);
  float lse = gm + logf(s);
  int tgt = targets[r];
  if (threadIdx.x == 0) nll[r] = lse - row[tgt];
  if (dlogits) {
    bf16* d = dlogits + r * ldd;
    for (int c = threadIdx.x * 4; c < V; c += 1024 * 4) {
      float4 x = *reinterpret_cast<const float4*>(row + c);
      float p0 = __expf(x.x - lse), p1 = __expf(x.y - lse), p2 = __expf(x.z - lse), p3 = __expf(x.w - lse);
      if (tgt >= c && tgt < c + 4) {
        int k = tgt - c;
        if (k == 0) p0 -= 1.f; else if (k == 1) p1 -= 1.f; else if (k == 2) p2 -= 1.f; else p3 -= 1.f;
      }
      uint2 o;
      o.x = pack_bf16x2(w * p0, w * p1);
      o.y = pack_bf16x2(w * p2, w * p3);
      *reinterpret_cast<uint2*>(d + c) = o;
    }
  }
}

// per-sample action MSE + gradient.  v,u [B, AD]; loss[b] = mean((v-u)^2); dv = gscale * 2 (v-u) / AD
__global__ void __launch_bounds__(128)
mse_fwd_bwd_kernel(const float* __restrict__ v, const float* __restrict__ u, float* __restrict__ loss,
                   float* __restrict__ dv, int AD, float gscale) {
  __shared__ float red[32];
  long b = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < AD; i += 128) {
    float d = v[b * AD + i] - u[b * AD + i];
    acc += d * d;
    if (dv) dv[b * AD + i] = gscale * 2.0f * d / AD;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[b] = acc / AD;
}

// out[0] (+)= alpha * sum_i x[i] * (w ? w[i] : 1)   — single CTA; n is small (rows / batch)
__global__ void __launch_bounds__(256)
weighted_sum_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out, long n,
                    float alpha, int accumulate) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long i = threadIdx.x; i < n; i += 256) acc += x[i] * (w ? w[i] : 1.0f);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + alpha * acc;
}

__global__ void sqrt_scalar_kernel(float* buf, int src, int dst) { buf[dst] = sqrtf(buf[src]); }

}  // namespace lapb

using namespace lapb;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int lapb200_ce_fwd_bwd(const float* logits, int64_t ld, const int32_t* targets, const float* weights, float* nll,
                       void* dlogits, int64_t ldd, int64_t R, int64_t V, lapb_stream_t s) {
  if (R == 0) return 0;
  LAPB_REQUIRE(V % 4 == 0 && ld % 4 == 0 && ldd % 4 == 0, "ce: V, ld, ldd must be multiples of 4");
  ce_fwd_bwd_kernel<<<(unsigned)R, 1024, 0, STREAM(s)>>>(logits, ld, targets, weights, nll, (bf16*)dlogits, ldd,
                                                        (int)V);
  LAPB_LAUNCH_OK("ce_fwd_bwd");
  return 0;
}

int lapb200_mse_fwd_bwd(const float* v, const float* u, float* loss, float* dv, int64_t B, int64_t AD, float gscale,
                        lapb_stream_t s) {
  mse_fwd_bwd_kernel<<<(unsigned)B, 128, 0, STREAM(s)>>>(v, u, loss, dv, (int)AD, gscale);
  LAPB_LAUNCH_OK("mse_fwd_bwd");
  return 0;
}

int lapb200_weighted_sum(const float* x, const float* w, float* out, int64_t n, float alpha, int64_t accumulate,
                         lapb_stream_t s) {
  weighted_sum_kernel<<<1, 256, 0, STREAM(s)>>>(x, w, out, n, alpha, (int)accumulate);
  LAPB_LAUNCH_OK("weighted_sum");
  return 0;
}

int lapb200_sqrt_scalar(float* buf, int64_t src, int64_t dst, lapb_stream_t s) {
  sqrt_scalar_kernel<<<1, 1, 0, STREAM(s)>>>(buf, (int)src, (int)dst);
  LAPB_LAUNCH_OK("sqrt_scalar");
  return 0;
}

}  // extern "C"
