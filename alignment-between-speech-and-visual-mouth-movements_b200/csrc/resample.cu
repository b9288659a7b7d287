// Sample-rate conversion in front of K1 (second half): FeatureExtractor.build_feature resamples every
// clip whose audio is not at cfg.sample_rate (misalignment_detection_train.py:202-204, librosa.resample) before the
// shift / MFCC stage.  librosa's default filter (soxr_hq) lives in an un-vendored C library and cannot be pinned here;
// this is a band-limited polyphase interpolator with a Kaiser-windowed sinc — the design librosa shipped as
// "kaiser_best" (resampy) before 0.10: 64 zero crossings, roll-off 0.9476, beta 14.77 — in the exact formulation of
// torchaudio.functional.resample(resampling_method="sinc_interp_kaiser"), which the oracle (oracle/resample_ref.py)
// restates and is tested against.
//
//   orig, new = orig_sr / g, new_sr / g (g = gcd);  base = min(orig, new) * rolloff;  width = ceil(zeros * orig / base)
//   h[i][k] = sinc(pi t) * kaiser(t) * base / orig,   t = clamp(((k - width) / orig - i / new) * base, +-zeros)
//   y[j * new + i] = sum_k h[i][k] * x[j * orig + k - width]          (x zero outside [0, n)),  len(y) = ceil(new * n / orig)
//
// One thread per output sample; the 2 * width + orig taps of a phase are contiguous in the table (L2-resident, 0.5 MB
// for 44.1 -> 16 kHz) and only the taps inside the window's support are visited.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>
#include "common.cuh"

struct avs_resample_plan {
  int orig, neu, width, taps;     // reduced rates, half-width in input samples, taps per phase = 2 * width + orig
  float* d_h = nullptr;           // [neu][taps]
  int* d_lo = nullptr;            // [neu] first tap with a non-zero weight
  int* d_hi = nullptr;            // [neu] one past the last
};

namespace avs {

__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ x, long long n_in, float* __restrict__ y, long long n_out, int n_signals,
                const float* __restrict__ h, const int* __restrict__ lo, const int* __restrict__ hi, int orig, int neu,
                int width, int taps) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= n_out * n_signals) return;
  const long long n = idx % n_out;
  const float* xs = x + (idx / n_out) * n_in;
  const long long j = n / neu;
  const int i = static_cast<int>(n - j * neu);
  const float* hp = h + static_cast<size_t>(i) * taps;
  const long long base = j * orig - width;           // x index of tap 0
  int k0 = lo[i], k1 = hi[i];
  if (base + k0 < 0) k0 = static_cast<int>(-base);
  if (base + k1 > n_in) k1 = static_cast<int>(n_in - base);
  // two interleaved accumulators in float, summed in tap order within each: the oracle accumulates in float64 and the
  // stated tolerance covers the difference
  float a0 = 0.f, a1 = 0.f;
  int k = k0;
  for (; k + 1 < k1; k += 2) {
    a0 = fmaf(__ldg(hp + k), __ldg(xs + base + k), a0);
    a1 = fmaf(__ldg(hp + k + 1), __ldg(xs + base + k + 1), a1);
  }
  if (k < k1) a0 = fmaf(__ldg(hp + k), __ldg(xs + base + k), a0);
  y[idx] = a0 + a1;
}

static double bessel_i0(double x) {  // power series, converges fast for the arguments used (<= 15)
  double sum = 1.0, term = 1.0;
  const double q = x * x / 4.0;
  for (int k = 1; k < 200; ++k) {
    term *= q / (static_cast<double>(k) * k);
    sum += term;
    if (term < 1e-17 * sum) break;
  }
  return sum;
}

}  // namespace avs

using namespace avs;

extern "C" int avs_resample_plan_create(int orig_sr, int target_sr, avs_resample_plan** out) {
  AVS_REQUIRE(out != nullptr, "null argument");
  AVS_REQUIRE(orig_sr > 0 && target_sr > 0, "sample rates must be positive");
  const int g = std::gcd(orig_sr, target_sr);
  avs_resample_plan* p = new avs_resample_plan();
  p->orig = orig_sr / g;
  p->neu = target_sr / g;
  const double zeros = 64.0, rolloff = 0.9475937167399596, beta = 14.769656459379492;
  const double base = std::min(p->orig, p->neu) * rolloff;
  p->width = static_cast<int>(std::ceil(zeros * p->orig / base));
  p->taps = 2 * p->width + p->orig;
  if (static_cast<double>(p->taps) * p->neu > 64e6) {
    set_error("resample %d -> %d Hz: the polyphase table would need %d x %d taps (rates with a small common divisor)", orig_sr,
              target_sr, p->neu, p->taps);
    delete p;
    return AVS_EINVAL;
  }
  std::vector<float> h(static_cast<size_t>(p->neu) * p->taps);
  std::vector<int> lo(p->neu), hi(p->neu);
  const double kPi = 3.14159265358979323846, i0b = bessel_i0(beta);
  for (int i = 0; i < p->neu; ++i) {
    int first = p->taps, last = 0;
    for (int k = 0; k < p->taps; ++k) {
      double t = (static_cast<double>(k - p->width) / p->orig - static_cast<double>(i) / p->neu) * base;
      t = std::max(-zeros, std::min(zeros, t));
      const double r = 1.0 - (t / zeros) * (t / zeros);
      const double win = bessel_i0(beta * std::sqrt(std::max(r, 0.0))) / i0b;
      const double a = t * kPi;
      const double v = (a == 0.0 ? 1.0 : std::sin(a) / a) * win * (base / p->orig);
      const float vf = static_cast<float>(v);
      h[static_cast<size_t>(i) * p->taps + k] = vf;
      if (std::fabs(t) < zeros) {  // inside the window's support (at the clamp the sinc is sin(64 pi) / (64 pi) ~ 0)
        first = std::min(first, k);
        last = std::max(last, k + 1);
      }
    }
    lo[i] = std::min(first, last);
    hi[i] = last;
  }
  if (cudaMalloc(reinterpret_cast<void**>(&p->d_h), h.size() * sizeof(float)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&p->d_lo), lo.size() * sizeof(int)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&p->d_hi), hi.size() * sizeof(int)) != cudaSuccess ||
      cudaMemcpy(p->d_h, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->d_lo, lo.data(), lo.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->d_hi, hi.data(), hi.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("resample_plan_create: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(p->d_h); cudaFree(p->d_lo); cudaFree(p->d_hi);
    delete p;
    return AVS_ECUDA;
  }
  *out = p;
  return AVS_OK;
}

extern "C" void avs_resample_plan_destroy(avs_resample_plan* p) {
  if (!p) return;
  cudaFree(p->d_h); cudaFree(p->d_lo); cudaFree(p->d_hi);
  delete p;
}

extern "C" long long avs_resample_out_len(const avs_resample_plan* p, long long n_in) {
  if (!p || n_in <= 0) return 0;
  return (static_cast<long long>(p->neu) * n_in + p->orig - 1) / p->orig;  // ceil(new * n / orig)
}

extern "C" int avs_resample(const avs_resample_plan* p, const float* in, long long n_in, int n_signals, float* out,
                            void* stream) {
  AVS_REQUIRE(p && in && out, "null argument");
  AVS_REQUIRE(n_signals >= 0 && n_in >= 0, "bad shape");
  const long long n_out = avs_resample_out_len(p, n_in);
  if (n_out == 0 || n_signals == 0) return AVS_OK;
  const long long total = n_out * n_signals;
  AVS_REQUIRE(total < (1LL << 40), "too many output samples");
  resample_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, n_in, out, n_out, n_signals, p->d_h, p->d_lo, p->d_hi, p->orig, p->neu, p->width, p->taps);
  AVS_LAUNCHED();
  return AVS_OK;
}
