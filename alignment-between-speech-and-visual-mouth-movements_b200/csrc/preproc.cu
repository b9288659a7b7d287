// Pre-processing prologue: the per-frame arithmetic of GridDataset.process_video
// (dataset.py:209-254) on the GPU, bit-exact with OpenCV's 8-bit paths:
//   BGR -> gray      cv2.cvtColor(COLOR_BGR2GRAY) : (B*3735 + G*19235 + R*9798 + 2^14) >> 15
//   crop             gray[int(h*0.6):, int(w*0.3):int(w*0.7)]   (whole frame if that region is empty)
//   resize           cv2.resize(., (100, 50)) INTER_LINEAR, 8-bit fixed point: 11-bit coefficients computed
//                    in float32, horizontal pass in int32, vertical pass
//                    (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
//   / 255.0          in double, rounded to float32 (256-entry table)
//   pad / truncate   to 75 frames (zeros)
// Exactness against cv2 holds when the cropped region has at least 50 rows (no vertical up-scaling);
// GRID videos (288 x 360 -> 116 x 144 crop) are in that regime.  Video DECODING stays on the host.
#include <cmath>
#include <vector>
#include "common.cuh"

struct avs_preproc {
  int h, w, channels, y0, x0, ch, cw;  // frame size, crop origin and crop size
  int* d_tab = nullptr;                // [3][100] x (index, a0, a1) then [3][50] y (index, b0, b1)
  float* d_lut = nullptr;              // [256] float(v / 255.0)
};

namespace avs {

constexpr int kOutW = AVS_W, kOutH = AVS_H, kOutT = AVS_T;

__device__ __forceinline__ int gray_at(const uint8_t* __restrict__ f, int w, int channels, int y, int x) {
  const uint8_t* p = f + (static_cast<size_t>(y) * w + x) * channels;
  if (channels == 1) return p[0];
  return (p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + (1 << 14)) >> 15;
}

// one thread per output pixel; lengths[b] (nullable) = valid frames of clip b, frames beyond it are zero
__global__ void __launch_bounds__(256)
preprocess_kernel(const uint8_t* __restrict__ frames, int n_in, const int32_t* __restrict__ lengths, avs_preproc pp,
                  float* __restrict__ out, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % kOutW);
  long long r = idx / kOutW;
  const int y = static_cast<int>(r % kOutH);
  r /= kOutH;
  const int t = static_cast<int>(r % kOutT);
  const long long b = r / kOutT;
  const int len = lengths ? min(lengths[b], n_in) : n_in;
  if (t >= len) {
    out[idx] = 0.f;
    return;
  }
  const int* tx = pp.d_tab;
  const int* ty = pp.d_tab + 3 * kOutW;
  const int xi = tx[x], a0 = tx[kOutW + x], a1 = tx[2 * kOutW + x];
  const int yi = ty[y], b0 = ty[kOutH + y], b1 = ty[2 * kOutH + y];
  const int xn = min(xi + 1, pp.cw - 1), yn = min(yi + 1, pp.ch - 1);
  const uint8_t* f = frames + (b * n_in + t) * static_cast<size_t>(pp.h) * pp.w * pp.channels;
  const int r0 = gray_at(f, pp.w, pp.channels, pp.y0 + yi, pp.x0 + xi) * a0 + gray_at(f, pp.w, pp.channels, pp.y0 + yi, pp.x0 + xn) * a1;
  const int r1 = gray_at(f, pp.w, pp.channels, pp.y0 + yn, pp.x0 + xi) * a0 + gray_at(f, pp.w, pp.channels, pp.y0 + yn, pp.x0 + xn) * a1;
  const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
  out[idx] = pp.d_lut[min(max(v, 0), 255)];
}

// cv::resize's linear coefficient table for one axis (float32 arithmetic, as OpenCV computes it)
static void linear_coeffs(int dn, int sn, std::vector<int>& tab) {
  tab.assign(3 * dn, 0);
  const double scale = static_cast<double>(sn) / dn;
  for (int d = 0; d < dn; ++d) {
    float f = static_cast<float>((d + 0.5) * scale - 0.5);
    int s = static_cast<int>(std::floor(f));
    f -= static_cast<float>(s);
    if (s < 0) s = 0, f = 0.f;
    if (s >= sn - 1) s = sn - 1, f = 0.f;
    tab[d] = s;
    tab[dn + d] = static_cast<int>(std::nearbyint((1.f - f) * 2048.f));
    tab[2 * dn + d] = static_cast<int>(std::nearbyint(f * 2048.f));
  }
}

}  // namespace avs

using namespace avs;

extern "C" int avs_preproc_create(int h, int w, int channels, avs_preproc** out) {
  AVS_REQUIRE(out && h > 0 && w > 0 && (channels == 1 || channels == 3), "frames must be [n, h, w, 1|3] uint8");
  avs_preproc* p = new avs_preproc();
  p->h = h; p->w = w; p->channels = channels;
  // dataset.py:216-217 — Python int() of a double product
  p->y0 = static_cast<int>(h * 0.6);
  p->x0 = static_cast<int>(w * 0.3);
  const int x1 = static_cast<int>(w * 0.7);
  p->ch = h - p->y0;
  p->cw = x1 - p->x0;
  if (p->ch <= 0 || p->cw <= 0) {  // :220-221 — empty region: use the full frame
    p->y0 = p->x0 = 0; p->ch = h; p->cw = w;
  }
  std::vector<int> tx, ty, tab;
  linear_coeffs(kOutW, p->cw, tx);
  linear_coeffs(kOutH, p->ch, ty);
  tab = tx;
  tab.insert(tab.end(), ty.begin(), ty.end());
  std::vector<float> lut(256);
  for (int v = 0; v < 256; ++v) lut[v] = static_cast<float>(v / 255.0);
  if (cudaMalloc(reinterpret_cast<void**>(&p->d_tab), tab.size() * sizeof(int)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&p->d_lut), 256 * sizeof(float)) != cudaSuccess ||
      cudaMemcpy(p->d_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->d_lut, lut.data(), 256 * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("preproc_create: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(p->d_tab); cudaFree(p->d_lut);
    delete p;
    return AVS_ECUDA;
  }
  *out = p;
  return AVS_OK;
}

extern "C" void avs_preproc_destroy(avs_preproc* p) {
  if (!p) return;
  cudaFree(p->d_tab);
  cudaFree(p->d_lut);
  delete p;
}

extern "C" int avs_preproc_crop(const avs_preproc* p, int* y0, int* x0, int* crop_h, int* crop_w) {
  AVS_REQUIRE(p && y0 && x0 && crop_h && crop_w, "null argument");
  *y0 = p->y0; *x0 = p->x0; *crop_h = p->ch; *crop_w = p->cw;
  return AVS_OK;
}

extern "C" int avs_preproc_run(const avs_preproc* p, const uint8_t* frames, int n_clips, int n_frames_in,
                               const int32_t* lengths, float* out, void* stream) {
  AVS_REQUIRE(p && frames && out, "null argument");
  AVS_REQUIRE(n_frames_in > 0, "no input frames");
  if (n_clips <= 0) return AVS_OK;
  const long long total = static_cast<long long>(n_clips) * kOutT * kOutH * kOutW;
  preprocess_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      frames, n_frames_in, lengths, *p, out, total);
  AVS_LAUNCHED();
  return AVS_OK;
}
