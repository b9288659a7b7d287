// fp32 CUDA-core GEMM used where fp32-exact accumulation matters more than tensor throughput
// (detector hidden layer, GRU projections in AVS_PREC_FP32):  C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]).
#pragma once
#include "common.cuh"
namespace avs {
// A row-major with leading dimension lda, B row-major [N, K] with ldb (torch Linear weight), C with ldc.
// Requires K % 8 == 0, lda % 4 == 0 and 16-byte aligned A; B rows of any alignment (scalar loads when ldb % 4 != 0).
int sgemm_nt(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M, int N,
             int K, cudaStream_t st);
int sgemm_nt_splitk(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int M, int N, int K,
                    int splits, float* partial, cudaStream_t st);
// C[M, N] = sum over the `splits` dense [M, N] slices of `partial`, in slice order, + bias (nullable)
int splitk_reduce(const float* partial, const float* bias, float* C, int M, int N, int splits, cudaStream_t st);
}  // namespace avs
