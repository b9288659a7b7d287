#include "sgemm.cuh"
namespace avs {

constexpr int BM = 128, BN = 128, BK = 8;

// BVEC: rows of B are 16-byte aligned (ldb % 4 == 0) and are staged with float4 loads; otherwise four scalar loads
// (the detector's W1 has ldb = 13824 + 2*n_mfcc: odd n_mfcc, e.g. the common 13, is not a multiple of 4).
__device__ __forceinline__ float4 ld4(const float* p, bool vec) {
  if (vec) return *reinterpret_cast<const float4*>(p);
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

template <bool BVEC>
__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N, int K) {
  // split-K: slice z covers k in [z*K, (z+1)*K) and writes its own [M, ldc] partial (bias must be null then)
  A += static_cast<size_t>(blockIdx.z) * K;
  B += static_cast<size_t>(blockIdx.z) * K;
  C += static_cast<size_t>(blockIdx.z) * M * ldc;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lr = tid >> 1, lk = (tid & 1) * 4;  // each thread stages one float4 of A and one of B per k-tile
  // 16 x 16 threads, 8 x 8 outputs each: rows ty*4 .. +3 and 64 + ty*4 .. +3, columns likewise — two 128-bit shared-memory
  // loads per operand and k instead of eight scalar ones (the kernel was LDS-issue bound: 16 LDS per 64 FMA)
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const bool a_ok = (m0 + lr) < M, b_ok = (n0 + lr) < N;
  const float* ap = A + static_cast<size_t>(m0 + lr) * lda + lk;
  const float* bp = B + static_cast<size_t>(n0 + lr) * ldb + lk;
  float4 ra = a_ok ? *reinterpret_cast<const float4*>(ap) : make_float4(0, 0, 0, 0);
  float4 rb = b_ok ? ld4(bp, BVEC) : make_float4(0, 0, 0, 0);
  const int nk = K / BK;  // K % 8 == 0 is enforced by the host wrapper via padding check
  int buf = 0;
  As[0][lk + 0][lr] = ra.x; As[0][lk + 1][lr] = ra.y; As[0][lk + 2][lr] = ra.z; As[0][lk + 3][lr] = ra.w;
  Bs[0][lk + 0][lr] = rb.x; Bs[0][lk + 1][lr] = rb.y; Bs[0][lk + 2][lr] = rb.z; Bs[0][lk + 3][lr] = rb.w;
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) {
      ra = a_ok ? *reinterpret_cast<const float4*>(ap + (kt + 1) * BK) : make_float4(0, 0, 0, 0);
      rb = b_ok ? ld4(bp + (kt + 1) * BK, BVEC) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]), a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]), b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      const int nb = buf ^ 1;
      As[nb][lk + 0][lr] = ra.x; As[nb][lk + 1][lr] = ra.y; As[nb][lk + 2][lr] = ra.z; As[nb][lk + 3][lr] = ra.w;
      Bs[nb][lk + 0][lr] = rb.x; Bs[nb][lk + 1][lr] = rb.y; Bs[nb][lk + 2][lr] = rb.z; Bs[nb][lk + 3][lr] = rb.w;
      __syncthreads();
      buf = nb;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
      if (n < N) C[static_cast<size_t>(m) * ldc + n] = acc[i][j] + (bias ? bias[n] : 0.f);
    }
  }
}

// Few outputs (one clip's hidden layer, the 39-class head of a clip or two): the 128 x 128 tiles above would be one to a
// few CTAs walking K alone (config 1: 128 us for the detector's hidden layer of ONE clip, 75 us for the class head).  Here
// every thread owns one output of one K slice and walks its rows of A and B with 128-bit loads — the SAME single fmaf
// chain in ascending k as the tiled kernel, so a result does not depend on which kernel (i.e. which batch size) computed it.
template <bool BVEC>
__global__ void __launch_bounds__(128)
sgemm_nt_skinny_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                       const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N, int K) {
  A += static_cast<size_t>(blockIdx.z) * K;
  B += static_cast<size_t>(blockIdx.z) * K;
  C += static_cast<size_t>(blockIdx.z) * M * ldc;
  const int idx = blockIdx.x * 128 + threadIdx.x;
  if (idx >= M * N) return;
  const int n = idx % N, m = idx / N;
  const float* ap = A + static_cast<size_t>(m) * lda;
  const float* bp = B + static_cast<size_t>(n) * ldb;
  float acc = 0.f;
  // 16 row segments of B in flight per thread (the kernel is a latency-bound stream of B: 28 MB for the detector's hidden
  // layer, read by 8192 threads), then the fmaf chain over them in k order; A comes from L1 (its rows are shared)
  constexpr int U = 16;
  for (int k0 = 0; k0 < K; k0 += 4 * U) {
    float4 b[U];
#pragma unroll
    for (int i = 0; i < U; ++i)
      if (k0 + 4 * i < K) b[i] = ld4(bp + k0 + 4 * i, BVEC);
#pragma unroll
    for (int i = 0; i < U; ++i)
      if (k0 + 4 * i < K) {
        const float4 a = *reinterpret_cast<const float4*>(ap + k0 + 4 * i);
        acc = fmaf(a.x, b[i].x, acc);
        acc = fmaf(a.y, b[i].y, acc);
        acc = fmaf(a.z, b[i].z, acc);
        acc = fmaf(a.w, b[i].w, acc);
      }
  }
  C[static_cast<size_t>(m) * ldc + n] = acc + (bias ? bias[n] : 0.f);
}
constexpr int kSkinnyOutputs = 16384;  // outputs x K slices up to which the one-thread-per-output kernel is used

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ C, int M, int N,
                     int splits) {
  const size_t n = static_cast<size_t>(M) * N;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[z * n + i];
    C[i] = s + (bias ? bias[i % N] : 0.f);
  }
}

// skinny-M variant: K is cut into `splits` slices (deterministic two-phase reduction through `partial`,
// which must hold splits * M * N floats); C is dense [M, N].
int sgemm_nt_splitk(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int M, int N, int K,
                    int splits, float* partial, cudaStream_t st) {
  AVS_REQUIRE(splits >= 1 && K % (splits * BK) == 0 && lda % 4 == 0, "bad split-K configuration");
  if (M <= 0 || N <= 0) return AVS_OK;
  if (splits == 1) return sgemm_nt(A, lda, B, ldb, bias, C, N, M, N, K, st);
  const bool bvec = ldb % 4 == 0 && reinterpret_cast<uintptr_t>(B) % 16 == 0;
  if (static_cast<long long>(M) * N * splits <= kSkinnyOutputs) {
    dim3 grid(cdiv(M * N, 128), 1, splits);
    if (bvec) sgemm_nt_skinny_kernel<true><<<grid, 128, 0, st>>>(A, lda, B, ldb, nullptr, partial, N, M, N, K / splits);
    else sgemm_nt_skinny_kernel<false><<<grid, 128, 0, st>>>(A, lda, B, ldb, nullptr, partial, N, M, N, K / splits);
  } else {
    dim3 grid(cdiv(N, BN), cdiv(M, BM), splits);
    if (bvec) sgemm_nt_kernel<true><<<grid, 256, 0, st>>>(A, lda, B, ldb, nullptr, partial, N, M, N, K / splits);
    else sgemm_nt_kernel<false><<<grid, 256, 0, st>>>(A, lda, B, ldb, nullptr, partial, N, M, N, K / splits);
  }
  AVS_LAUNCHED();
  return splitk_reduce(partial, bias, C, M, N, splits, st);
}

int splitk_reduce(const float* partial, const float* bias, float* C, int M, int N, int splits, cudaStream_t st) {
  splitk_reduce_kernel<<<cdiv(M * N, 1024), 256, 0, st>>>(partial, bias, C, M, N, splits);
  AVS_LAUNCHED();
  return AVS_OK;
}

int sgemm_nt(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M, int N,
             int K, cudaStream_t st) {
  AVS_REQUIRE(K % BK == 0 && lda % 4 == 0, "sgemm_nt needs K % 8 == 0 and 16-byte aligned rows of A");
  if (M <= 0 || N <= 0) return AVS_OK;
  const bool bvec = ldb % 4 == 0 && reinterpret_cast<uintptr_t>(B) % 16 == 0;
  if (static_cast<long long>(M) * N <= kSkinnyOutputs) {
    if (bvec) sgemm_nt_skinny_kernel<true><<<cdiv(M * N, 128), 128, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K);
    else sgemm_nt_skinny_kernel<false><<<cdiv(M * N, 128), 128, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K);
  } else {
    dim3 grid(cdiv(N, BN), cdiv(M, BM));
    if (bvec) sgemm_nt_kernel<true><<<grid, 256, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K);
    else sgemm_nt_kernel<false><<<grid, 256, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K);
  }
  AVS_LAUNCHED();
  return AVS_OK;
}

}  // namespace avs
