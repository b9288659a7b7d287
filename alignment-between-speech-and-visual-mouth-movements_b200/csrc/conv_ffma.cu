// STCNN layer on CUDA cores, fp32 throughout (AVS_PREC_FP32): Conv3d(k = 3 x KH x KW, "same" padding)
// + bias + ReLU + MaxPool3d((1,2,2)) fused (model.py:67-76).  This is the bring-up / cross-check
// path and the fp32-exact arithmetic mode; the throughput path is conv_umma.cu.
//
// One CTA computes, for one (clip, t), a 16 x 16 tile of POOLED outputs (32 x 32 conv outputs) for
// CO_T = 8 output channels: each thread owns one pooled position = a 2 x 2 conv quad x 8 channels.
// Input patches (3 time planes x (32 + KH - 1) x (32 + KW - 1)) and the weights of the current input
// channel are staged in shared memory.
#include <type_traits>
#include "stcnn.cuh"

namespace avs {

template <int KH, int KW>
__global__ void __launch_bounds__(256)
conv_pool_ffma_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ bias,
                      float* __restrict__ out, int Cin, int Cout, int T, int H, int W,
                      long long o_sb, long long o_sc, long long o_st) {
  constexpr int CO_T = 8, PT = 16, CT = 32;
  constexpr int PH = CT + KH - 1, PW = CT + KW - 1, PWP = PW + 1;
  constexpr int PADH = KH / 2, PADW = KW / 2;
  __shared__ float s_in[3][PH][PWP];
  __shared__ float s_w[3 * KH * KW][CO_T];
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_w = (Wo + PT - 1) / PT;
  const int tile_h = blockIdx.x / tiles_w, tile_w = blockIdx.x % tiles_w;
  const int t = blockIdx.y % T, b = blockIdx.y / T;
  const int co0 = blockIdx.z * CO_T;
  const int tid = threadIdx.x, ty = tid / PT, tx = tid % PT;
  const int h0 = tile_h * CT, w0 = tile_w * CT;  // conv-output origin of this tile

  float acc[2][2][CO_T];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < CO_T; ++c) acc[i][j][c] = 0.f;

  for (int ci = 0; ci < Cin; ++ci) {
    __syncthreads();
    for (int i = tid; i < 3 * PH * PW; i += 256) {
      const int kd = i / (PH * PW), r = (i / PW) % PH, c = i % PW;
      const int tt = t + kd - 1, hh = h0 + r - PADH, ww = w0 + c - PADW;
      float v = 0.f;
      if (tt >= 0 && tt < T && hh >= 0 && hh < H && ww >= 0 && ww < W)
        v = in[(((static_cast<size_t>(b) * Cin + ci) * T + tt) * H + hh) * W + ww];
      s_in[kd][r][c] = v;
    }
    for (int i = tid; i < 3 * KH * KW * CO_T; i += 256) {
      const int tap = i / CO_T, c = i % CO_T;
      s_w[tap][c] = (co0 + c < Cout) ? wgt[(static_cast<size_t>(co0 + c) * Cin + ci) * (3 * KH * KW) + tap] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
      for (int r = 0; r < KH + 1; ++r) {  // input row r of the quad's (KH+1)-row window
        float x[KW + 1];
#pragma unroll
        for (int c = 0; c < KW + 1; ++c) x[c] = s_in[kd][2 * ty + r][2 * tx + c];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int kh = r - i;
          if (kh < 0 || kh >= KH) continue;
#pragma unroll
          for (int kw = 0; kw < KW; ++kw) {
            const float4 wa = *reinterpret_cast<const float4*>(&s_w[(kd * KH + kh) * KW + kw][0]);
            const float4 wb = *reinterpret_cast<const float4*>(&s_w[(kd * KH + kh) * KW + kw][4]);
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int c = 0; c < CO_T; ++c) acc[i][j][c] = fmaf(x[kw + j], wv[c], acc[i][j][c]);
          }
        }
      }
    }
  }
  const int ho = tile_h * PT + ty, wo = tile_w * PT + tx;
  if (ho < Ho && wo < Wo) {
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
      if (co0 + c >= Cout) continue;
      const float m = fmaxf(fmaxf(acc[0][0][c], acc[0][1][c]), fmaxf(acc[1][0][c], acc[1][1][c]));
      out[b * o_sb + (co0 + c) * o_sc + t * o_st + static_cast<long long>(ho) * Wo + wo] = fmaxf(m + bias[co0 + c], 0.f);
    }
  }
}

int conv_pool_ffma(const float* in, const float* w, const float* bias, float* out, int B, int Cin, int Cout, int T,
                   int H, int W, int KH, int KW, long long o_sb, long long o_sc, long long o_st, cudaStream_t st) {
  const int Ho = H / 2, Wo = W / 2;
  dim3 grid(cdiv(Ho, 16) * cdiv(Wo, 16), B * T, cdiv(Cout, 8));
  ProfScope ps(Cin == 1 ? PROF_CONV1 : (Cout == 64 ? PROF_CONV2 : PROF_CONV3), st);
  AVS_REQUIRE(static_cast<long long>(B) * T <= 65535, "FFMA conv path: n_clips * T must be <= 65535 per call");
  if (KH == 5 && KW == 5)
    conv_pool_ffma_kernel<5, 5><<<grid, 256, 0, st>>>(in, w, bias, out, Cin, Cout, T, H, W, o_sb, o_sc, o_st);
  else if (KH == 3 && KW == 3)
    conv_pool_ffma_kernel<3, 3><<<grid, 256, 0, st>>>(in, w, bias, out, Cin, Cout, T, H, W, o_sb, o_sc, o_st);
  else {
    set_error("unsupported kernel size %dx%d", KH, KW);
    return AVS_EINVAL;
  }
  AVS_LAUNCHED();
  return AVS_OK;
}

// visual statistics (misalignment_detection_train.py:165): emb [B, T, F] -> [B, 2F] = [mean_t, unbiased std_t]
template <int T>
__global__ void __launch_bounds__(128)
vstats_kernel(const float* __restrict__ emb, float* __restrict__ out, int F) {
  const int f = blockIdx.x * 128 + threadIdx.x, b = blockIdx.y;
  if (f >= F) return;
  const float* e = emb + static_cast<size_t>(b) * T * F + f;
  float v[T];
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    v[t] = e[static_cast<size_t>(t) * F];
    s += v[t];
  }
  const float mean = s / T;
  float ss = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float d = v[t] - mean;
    ss = fmaf(d, d, ss);
  }
  out[static_cast<size_t>(b) * 2 * F + f] = mean;
  out[static_cast<size_t>(b) * 2 * F + F + f] = sqrtf(ss / (T - 1));
}

// Fused-statistics path (conv3 epilogue mode 2): fold the per-(clip, part) partial sums (sum_t x, sum_t x^2) into
// [mean_t, unbiased std_t] (misalignment_detection_train.py:165) and leave the scratch zeroed for the next launch.
// Parts are added in part order, so the result does not depend on which CTA finished first.  The final arithmetic is
// f64: sum (x - mean)^2 = sum x^2 - (sum x)^2 / T.  With f64 partials (fp32-grade kind) that is exact to ~1e-16 of
// sum x^2; with f32 partials (bf16 kind) the partial sums carry ~1e-7 relative rounding, i.e. a relative error of about
// 1e-7 * (1 + mean^2 / var) on the variance — below the bf16 kind's own error (2^-9 per operand) for mean / std < 100.
template <typename StatT>
__global__ void __launch_bounds__(256)
vstats_finish_kernel(StatT* __restrict__ stat, int parts, float* __restrict__ out, int F, int T) {
  const int f = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (f >= F) return;
  using Pair = typename std::conditional<sizeof(StatT) == 8, double2, float2>::type;
  Pair* base = reinterpret_cast<Pair*>(stat) + static_cast<size_t>(b) * parts * F + f;
  Pair v[4];
  double s = 0.0, ss = 0.0;
  for (int p0 = 0; p0 < parts; p0 += 4) {  // loads of four parts in flight, then their zeroing stores
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (p0 + j < parts) v[j] = base[static_cast<size_t>(p0 + j) * F];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (p0 + j < parts) {
        s += static_cast<double>(v[j].x);
        ss += static_cast<double>(v[j].y);
        Pair z; z.x = 0; z.y = 0;
        base[static_cast<size_t>(p0 + j) * F] = z;
      }
  }
  const double mean = s / T;
  const double var = fmax(ss - s * mean, 0.0) / (T - 1);
  out[static_cast<size_t>(b) * 2 * F + f] = static_cast<float>(mean);
  out[static_cast<size_t>(b) * 2 * F + F + f] = static_cast<float>(sqrt(var));
}

int vstat_parts(int n_clips, int items_per_clip, int n_sms) {
  if (n_clips <= 0 || items_per_clip <= 0) return 1;
  const long long items = static_cast<long long>(n_clips) * items_per_clip;
  const long long grid = items < n_sms ? items : n_sms;
  const long long span_min = items / grid;                       // spans are floor or ceil of items / grid
  long long parts = (items_per_clip + span_min - 1) / span_min + 1;
  if (parts > items_per_clip) parts = items_per_clip;
  if (parts > grid) parts = grid;
  return static_cast<int>(parts);
}

size_t vstat_scratch_bytes(int cap_clips, int items_per_clip, int n_sms, bool f64) {
  size_t worst = 0;  // a call may bring any number of clips up to the capacity: clips x parts peaks for small batches
  for (int b = 1; b <= cap_clips; ++b) {
    const size_t n = static_cast<size_t>(b) * vstat_parts(b, items_per_clip, n_sms);
    if (n > worst) worst = n;
  }
  return worst * 2 * AVS_EMB * (f64 ? sizeof(double) : sizeof(float));
}

int vstats_finish(void* stat, bool f64, int parts, float* out, int B, cudaStream_t st) {
  ProfScope ps(PROF_VSTATS, st);
  const size_t clip_bytes = static_cast<size_t>(parts) * 2 * AVS_EMB * (f64 ? sizeof(double) : sizeof(float));
  for (int b0 = 0; b0 < B; b0 += 32768) {
    const int nb = B - b0 < 32768 ? B - b0 : 32768;
    const dim3 grid(cdiv(AVS_EMB, 256), nb);
    void* sp = static_cast<uint8_t*>(stat) + b0 * clip_bytes;
    float* op = out + static_cast<size_t>(b0) * 2 * AVS_EMB;
    if (f64) vstats_finish_kernel<double><<<grid, 256, 0, st>>>(static_cast<double*>(sp), parts, op, AVS_EMB, AVS_T);
    else vstats_finish_kernel<float><<<grid, 256, 0, st>>>(static_cast<float*>(sp), parts, op, AVS_EMB, AVS_T);
    AVS_LAUNCHED();
  }
  return AVS_OK;
}

int vstats(const float* emb, float* out, int B, int F, cudaStream_t st) {
  if (B <= 0) return AVS_OK;
  ProfScope ps(PROF_VSTATS, st);
  for (int b0 = 0; b0 < B; b0 += 32768) {
    const int nb = B - b0 < 32768 ? B - b0 : 32768;
    vstats_kernel<AVS_T><<<dim3(cdiv(F, 128), nb), 128, 0, st>>>(emb + static_cast<size_t>(b0) * AVS_T * F,
                                                                 out + static_cast<size_t>(b0) * 2 * F, F);
    AVS_LAUNCHED();
  }
  return AVS_OK;
}

}  // namespace avs
