// Host-side helpers shared by the launchers in liblapb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace lapb {

char* error_buffer();  // defined in api.cu (static 512-byte buffer)

inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define LAPB_CUDA_OK(expr)                                                                      \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return lapb::set_error((int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                             __FILE__, __LINE__);                                               \
  } while (0)

#define LAPB_REQUIRE(cond, ...)                         \
  do {                                                  \
    if (!(cond)) return lapb::set_error(-1, __VA_ARGS__); \
  } while (0)

// Launch-status check that does not synchronize.
#define LAPB_LAUNCH_OK(name)                                                                   \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return lapb::set_error((int)_e, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

int num_sms();  // cached SM count of the current device (api.cu)

inline unsigned int cdiv(long a, long b) { return (unsigned int)((a + b - 1) / b); }

// LAPB_PDL=1: kernels that carry griddepcontrol.wait are launched with the programmatic-stream-serialization attribute
// (their prologue overlaps the tail of the preceding kernel).  Read once per process.
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LAPB_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

}  // namespace lapb
