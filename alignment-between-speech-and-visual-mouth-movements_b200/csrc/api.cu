// Library-wide pieces of the C-ABI: version, error string, device check, launch counter.
#include <stdarg.h>
#include "common.cuh"
#include "stcnn.cuh"

namespace avs {
static thread_local char g_err[512] = "";
long long g_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- per-kernel timing
int g_prof_on = 0;
namespace {
constexpr int kProfRing = 8192;
struct ProfState {
  cudaEvent_t ev[kProfRing][2] = {};
  int used = 0, made = 0;   // events [0, made) exist
} g_prof[PROF_NSLOTS];
}
void prof_begin(int slot, cudaStream_t st) {
  ProfState& p = g_prof[slot];
  if (p.used >= kProfRing) return;
  if (p.used >= p.made) {
    cudaEventCreate(&p.ev[p.made][0]);
    cudaEventCreate(&p.ev[p.made][1]);
    ++p.made;
  }
  cudaEventRecord(p.ev[p.used][0], st);
}
void prof_end(int slot, cudaStream_t st) {
  ProfState& p = g_prof[slot];
  if (p.used >= kProfRing || p.used >= p.made) return;
  cudaEventRecord(p.ev[p.used][1], st);
  ++p.used;
}
}  // namespace avs

extern "C" void avs_prof_enable(int on) { avs::g_prof_on = on; }
extern "C" void avs_prof_reset(void) {
  for (int s = 0; s < avs::PROF_NSLOTS; ++s) avs::g_prof[s].used = 0;
}
extern "C" int avs_prof_read(int slot, double* total_ms, int* count) {
  if (slot < 0 || slot >= avs::PROF_NSLOTS || !total_ms || !count) return AVS_EINVAL;
  auto& p = avs::g_prof[slot];
  double t = 0;
  for (int i = 0; i < p.used; ++i) {
    AVS_CUDA(cudaEventSynchronize(p.ev[i][1]));
    float ms = 0;
    AVS_CUDA(cudaEventElapsedTime(&ms, p.ev[i][0], p.ev[i][1]));
    t += ms;
  }
  *total_ms = t;
  *count = p.used;
  return AVS_OK;
}

#ifdef AVS_EXPERIMENTS
namespace avs { extern int g_conv_dbg; }
extern "C" void avs_debug_set(int flags) { avs::g_conv_dbg = flags; }
// begin / end of the idx-th profiled launch of a slot, in ms since the first profiled pack launch (tools/sweep_timeline.py)
extern "C" int avs_prof_read_span(int slot, int idx, double* begin_ms, double* end_ms) {
  if (slot < 0 || slot >= avs::PROF_NSLOTS || !begin_ms || !end_ms) return AVS_EINVAL;
  auto& p = avs::g_prof[slot];
  auto& base = avs::g_prof[0];
  if (idx < 0 || idx >= p.used || base.used < 1) return AVS_EINVAL;
  AVS_CUDA(cudaEventSynchronize(p.ev[idx][1]));
  float b = 0, e = 0;
  AVS_CUDA(cudaEventElapsedTime(&b, base.ev[0][0], p.ev[idx][0]));
  AVS_CUDA(cudaEventElapsedTime(&e, base.ev[0][0], p.ev[idx][1]));
  *begin_ms = b;
  *end_ms = e;
  return AVS_OK;
}
#endif

extern "C" int avs_conv_item_span(int n_clips, int n_steps, int n_tiles, int tiles_per_item, int n_ctas, int cta, int* first,
                                  int* last) {
  AVS_REQUIRE(first && last, "null argument");
  AVS_REQUIRE(n_clips >= 0 && n_steps > 0 && n_tiles > 0 && tiles_per_item > 0 && n_ctas > 0 && cta >= 0 && cta < n_ctas, "bad argument");
  avs::conv_item_span(n_clips, n_steps, n_tiles, tiles_per_item, n_ctas, cta, first, last);
  return AVS_OK;
}

extern "C" int avs_version(void) { return AVS_VERSION; }
static const char kSourceHash[] =
#include "build_src_hash.inc"
    ;
extern "C" const char* avs_source_hash(void) { return kSourceHash; }
extern "C" const char* avs_last_error_string(void) { return avs::g_err; }
extern "C" long long avs_launch_count(void) { return avs::g_launches; }

extern "C" int avs_device_check(int device) {
  cudaDeviceProp prop;  // (callers cache the verdict per device: cudaGetDeviceProperties is slow)
  AVS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10 || prop.minor != 0) {  // the cubin is sm_100a: no other 10.x part can run it
    avs::set_error("device %d is sm_%d%d; libavsync_b200 is built for sm_100a only and has no fallback", device,
                   prop.major, prop.minor);
    return AVS_EARCH;
  }
  return AVS_OK;
}
