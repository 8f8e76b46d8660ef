// Image side of preprocess_observation on the device (src/lap/models/model_adapter.py:83-181), ahead of the patchify kernel:
//   image_resize_pad  model_adapter.py:113-116 -> OP/shared/image_tools.py:11-52 (resize_with_pad: aspect-preserving linear
//                     resize with the antialiasing triangle filter of jax.image.resize, clip / round, centred padding) fused
//                     with the uint8 -> [-1, 1] conversion of Observation.from_dict (OP/models/model.py:116-118)
//   image_augment     model_adapter.py:118-151: crop 95 % -> resize -> rotate as ONE bilinear resampling, then colour jitter
//                     (brightness / contrast / saturation), with every random quantity an explicit per-sample parameter
// Both are gather kernels over tiny inputs (a 32-sample batch of 224 x 224 x 3 images is 19 MB in fp32): one thread per
// output pixel, three channels per thread, HBM-bound and far below a microsecond-scale budget per image.
#include "../../include/lapb200.h"
#include "common.cuh"
#include "host_util.h"
#include <algorithm>

namespace lapb {

template <bool U8>
__device__ __forceinline__ float img_load(const void* src, long idx) {
  if (U8) return (float)reinterpret_cast<const uint8_t*>(src)[idx];
  return reinterpret_cast<const float*>(src)[idx];
}

// dst[b, y, x, :] for the padded [Hout, Wout] frame; the resized region is rows [ph0, ph0 + rh), columns [pw0, pw0 + rw).
// Separable filter given as sparse rows: output row p uses input rows ystart[p] .. ystart[p] + ytaps - 1 with weights
// yw[p * ytaps + t] (zero beyond the support), likewise for columns.
template <bool U8>
__global__ void __launch_bounds__(256)
image_resize_pad_kernel(const void* __restrict__ src, float* __restrict__ dst, int B, int Hin, int Win, int Hout, int Wout,
                        int rh, int rw, int ph0, int pw0, const int* __restrict__ ystart, const float* __restrict__ yw,
                        int ytaps, const int* __restrict__ xstart, const float* __restrict__ xw, int xtaps) {
  const long n = (long)B * Hout * Wout;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long)gridDim.x * blockDim.x) {
    const int x = (int)(o % Wout), y = (int)((o / Wout) % Hout), b = (int)(o / ((long)Wout * Hout));
    const int p = y - ph0, q = x - pw0;
    float r0 = -1.f, r1 = -1.f, r2 = -1.f;  // padding: 0 for uint8 == -1 after u8 / 255 * 2 - 1; -1 for float images
    if (p >= 0 && p < rh && q >= 0 && q < rw) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      const int ys = ystart[p], xs = xstart[q];
      for (int ty = 0; ty < ytaps; ++ty) {
        const int iy = ys + ty;
        const float wy = yw[p * ytaps + ty];
        if (iy < 0 || iy >= Hin || wy == 0.f) continue;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        const long rowbase = ((long)b * Hin + iy) * Win;
        for (int tx = 0; tx < xtaps; ++tx) {
          const int ix = xs + tx;
          const float wx = xw[q * xtaps + tx];
          if (ix < 0 || ix >= Win || wx == 0.f) continue;
          const long idx = (rowbase + ix) * 3;
          c0 += wx * img_load<U8>(src, idx);
          c1 += wx * img_load<U8>(src, idx + 1);
          c2 += wx * img_load<U8>(src, idx + 2);
        }
        a0 += wy * c0;
        a1 += wy * c1;
        a2 += wy * c2;
      }
      if (U8) {  // round half to even like jnp.round, clip, then Observation.from_dict's u8 / 255 * 2 - 1
        r0 = fminf(fmaxf(rintf(a0), 0.f), 255.f) / 255.0f * 2.0f - 1.0f;
        r1 = fminf(fmaxf(rintf(a1), 0.f), 255.f) / 255.0f * 2.0f - 1.0f;
        r2 = fminf(fmaxf(rintf(a2), 0.f), 255.f) / 255.0f * 2.0f - 1.0f;
      } else {
        r0 = fminf(fmaxf(a0, -1.f), 1.f);
        r1 = fminf(fmaxf(a1, -1.f), 1.f);
        r2 = fminf(fmaxf(a2, -1.f), 1.f);
      }
    }
    float* d = dst + o * 3;
    d[0] = r0; d[1] = r1; d[2] = r2;
  }
}

// params[b] = (crop_y, crop_x, angle_deg, brightness, contrast, saturation, skip, -)  — see oracle/image_oracle.py
template <bool U8>
__global__ void __launch_bounds__(256)
image_augment_kernel(const void* __restrict__ src, float* __restrict__ dst, int B, int H, int W, int ch, int cw,
                     const float* __restrict__ params) {
  const long n = (long)B * H * W;
  const float cy0 = (H - 1) * 0.5f, cx0 = (W - 1) * 0.5f;
  const float sy_scale = (float)ch / (float)H, sx_scale = (float)cw / (float)W;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long)gridDim.x * blockDim.x) {
    const int x = (int)(o % W), y = (int)((o / W) % H), b = (int)(o / ((long)W * H));
    const float* pr = params + (long)b * 8;
    float* d = dst + o * 3;
    auto pix = [&](int iy, int ix, int c) -> float {  // input pixel in [0, 1]
      const float v = img_load<U8>(src, (((long)b * H + iy) * W + ix) * 3 + c);
      return U8 ? (v / 255.0f * 2.0f - 1.0f) * 0.5f + 0.5f : v * 0.5f + 0.5f;
    };
    if (pr[6] > 0.5f) {  // VQA sample: untouched
      for (int c = 0; c < 3; ++c) {
        const float v = img_load<U8>(src, o * 3 + c);
        d[c] = U8 ? v / 255.0f * 2.0f - 1.0f : v;
      }
      continue;
    }
    float sn, cs;
    sincosf(pr[2] * 0.017453292519943295f, &sn, &cs);
    const float dy = (float)y - cy0, dx = (float)x - cx0;
    const float ry = cs * dy - sn * dx + cy0, rx = sn * dy + cs * dx + cx0;
    const float sy = (ry + 0.5f) * sy_scale - 0.5f + pr[0];
    const float sx = (rx + 0.5f) * sx_scale - 0.5f + pr[1];
    const float fy0 = floorf(sy), fx0 = floorf(sx);
    const float fy = sy - fy0, fx = sx - fx0;
    const int y0 = (int)fy0, x0 = (int)fx0;
    float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int oy = 0; oy < 2; ++oy) {
#pragma unroll
      for (int ox = 0; ox < 2; ++ox) {
        const int iy = y0 + oy, ix = x0 + ox;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        const float w = (oy ? fy : 1.0f - fy) * (ox ? fx : 1.0f - fx);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] += pix(iy, ix, c) * w;
      }
    }
    const float br = pr[3], co = pr[4], sa = pr[5];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v[c] = br < 0.f ? v[c] * (1.0f + br) : v[c] * (1.0f - br) + br;
      v[c] = (v[c] - 0.5f) * (1.0f + co) + 0.5f;
    }
    const float g = v[0] * 0.299f + v[1] * 0.587f + v[2] * 0.114f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float s = g + (v[c] - g) * (1.0f + sa);
      d[c] = fminf(fmaxf(s, 0.f), 1.f) * 2.0f - 1.0f;
    }
  }
}

}  // namespace lapb

using namespace lapb;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int lapb200_image_resize_pad(const void* src, int64_t src_is_u8, float* dst, int64_t B, int64_t Hin, int64_t Win,
                             int64_t Hout, int64_t Wout, int64_t rh, int64_t rw, int64_t ph0, int64_t pw0,
                             const int32_t* ystart, const float* yw, int64_t ytaps, const int32_t* xstart, const float* xw,
                             int64_t xtaps, lapb_stream_t s) {
  LAPB_REQUIRE(B >= 1 && rh >= 1 && rw >= 1 && ph0 >= 0 && pw0 >= 0 && ph0 + rh <= Hout && pw0 + rw <= Wout,
               "image_resize_pad: the resized region must lie inside the output frame");
  const long n = B * Hout * Wout;
  const int grid = (int)std::min<long>((n + 255) / 256, 148L * 16);
  if (src_is_u8)
    image_resize_pad_kernel<true><<<grid, 256, 0, STREAM(s)>>>(src, dst, (int)B, (int)Hin, (int)Win, (int)Hout, (int)Wout,
                                                               (int)rh, (int)rw, (int)ph0, (int)pw0, ystart, yw, (int)ytaps,
                                                               xstart, xw, (int)xtaps);
  else
    image_resize_pad_kernel<false><<<grid, 256, 0, STREAM(s)>>>(src, dst, (int)B, (int)Hin, (int)Win, (int)Hout, (int)Wout,
                                                                (int)rh, (int)rw, (int)ph0, (int)pw0, ystart, yw, (int)ytaps,
                                                                xstart, xw, (int)xtaps);
  LAPB_LAUNCH_OK("image_resize_pad");
  return 0;
}

int lapb200_image_augment(const void* src, int64_t src_is_u8, float* dst, int64_t B, int64_t H, int64_t W,
                          const float* params, lapb_stream_t s) {
  LAPB_REQUIRE(B >= 1 && H >= 2 && W >= 2 && src != (const void*)dst, "image_augment: bad sizes / in-place call");
  const int ch = (int)(H * 0.95), cw = (int)(W * 0.95);
  const long n = B * H * W;
  const int grid = (int)std::min<long>((n + 255) / 256, 148L * 16);
  if (src_is_u8)
    image_augment_kernel<true><<<grid, 256, 0, STREAM(s)>>>(src, dst, (int)B, (int)H, (int)W, ch, cw, params);
  else
    image_augment_kernel<false><<<grid, 256, 0, STREAM(s)>>>(src, dst, (int)B, (int)H, (int)W, ch, cw, params);
  LAPB_LAUNCH_OK("image_augment");
  return 0;
}

}  // extern "C"
