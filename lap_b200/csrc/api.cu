// Library-level entry points of liblapb200.so: error reporting, device check.
#include "../../include/lapb200.h"
#include "host_util.h"

namespace lapb {

static char g_error[512] = "ok";
char* error_buffer() { return g_error; }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace lapb

extern "C" const char* lapb200_last_error(void) { return lapb::error_buffer(); }

extern "C" int lapb200_version(void) { return 100; }

extern "C" int lapb200_check_device(void) {
  int dev = 0;
  LAPB_CUDA_OK(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  LAPB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  LAPB_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  LAPB_REQUIRE(major == 10 && minor == 0, "lapb200 requires an sm_100 (B200) device, found sm_%d%d; no fallback",
               major, minor);
  return 0;
}
