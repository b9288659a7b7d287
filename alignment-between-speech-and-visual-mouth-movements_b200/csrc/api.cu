// Library-wide pieces of the C-ABI: version, error string, device check, launch counter.
#include <stdarg.h>
#include "common.cuh"

namespace avs {
static thread_local char g_err[512] = "";
long long g_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace avs

extern "C" int avs_version(void) { return AVS_VERSION; }
extern "C" const char* avs_last_error_string(void) { return avs::g_err; }
extern "C" long long avs_launch_count(void) { return avs::g_launches; }

extern "C" int avs_device_check(int device) {
  cudaDeviceProp prop;
  AVS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    avs::set_error("device %d is sm_%d%d; libavsync_b200 is built for sm_100a only and has no fallback", device,
                   prop.major, prop.minor);
    return AVS_EARCH;
  }
  return AVS_OK;
}
