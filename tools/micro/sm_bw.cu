// Microbenchmark: how fast can a FEW SMs stream weights?  16 CTAs (one per SM) x 1024 threads read a buffer with
// 16-byte ld.global.nc loads, U loads in flight per thread; variants: cold (DRAM), L2-resident, and DRAM with an L2
// prefetch running one slice ahead.  Decides whether a 16-CTA cluster-resident denoise kernel can be bandwidth-bound.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
template <int U, bool PREFETCH>
__global__ void __launch_bounds__(1024, 1) stream_kernel(const uint4* __restrict__ buf, size_t n_per_cta, uint32_t* sink) {
  const uint4* p = buf + (size_t)blockIdx.x * n_per_cta;
  uint32_t acc = 0;
  const size_t step = 1024 * U;
  for (size_t i = threadIdx.x; i < n_per_cta; i += step) {
    if (PREFETCH) {  // prefetch the slice 64 iterations (1 MB at U=4) ahead into L2
      size_t j = i + 64 * step;
      if (j < n_per_cta && (threadIdx.x & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + j));
    }
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = (i + u * 1024 < n_per_cta) ? ld_stream(p + i + u * 1024) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}
template <int U, bool PF>
float run(const uint4* buf, size_t n_per_cta, int ctas, uint32_t* sink, bool flush, void* fl, size_t flbytes) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < 5; ++it) {
    if (flush) cudaMemsetAsync(fl, it, flbytes);
    cudaEventRecord(e0);
    stream_kernel<U, PF><<<ctas, 1024>>>(buf, n_per_cta, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}
int main() {
  const int ctas = 16;
  const size_t bytes_per_cta = 8u << 20;  // 8 MB per CTA: 128 MB total > L2 for the cold runs is not needed (flush instead)
  const size_t n_per_cta = bytes_per_cta / 16;
  uint4* buf; uint32_t* sink; void* fl; const size_t flbytes = 512u << 20;
  cudaMalloc(&buf, bytes_per_cta * 148); cudaMalloc(&sink, 4); cudaMalloc(&fl, flbytes);
  cudaMemset(buf, 1, bytes_per_cta * 148);
  for (int n : {16, 32, 148}) {
    float a = run<4, false>(buf, n_per_cta, n, sink, true, fl, flbytes);
    float b = run<8, false>(buf, n_per_cta, n, sink, true, fl, flbytes);
    float c = run<4, true>(buf, n_per_cta, n, sink, true, fl, flbytes);
    size_t small = (2u << 20) / 16;  // 2 MB per CTA: L2-resident on repeat
    run<4, false>(buf, small, n, sink, false, fl, flbytes);
    float d = run<4, false>(buf, small, n, sink, false, fl, flbytes);
    float e = run<8, false>(buf, small, n, sink, false, fl, flbytes);
    printf("ctas %3d | cold U4 %.1f GB/s/SM  cold U8 %.1f  cold+prefetch U4 %.1f | L2-resident U4 %.1f  U8 %.1f GB/s/SM\n", n,
           bytes_per_cta / a / 1e6, bytes_per_cta / b / 1e6, bytes_per_cta / c / 1e6, (2u << 20) / d / 1e6, (2u << 20) / e / 1e6);
  }
  return 0;
}
