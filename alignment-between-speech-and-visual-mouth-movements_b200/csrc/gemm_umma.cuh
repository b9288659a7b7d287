// tcgen05 hi/lo-split GEMM used by the Bi-GRU head (gemm_umma.cu)
#pragma once
#include "common.cuh"
namespace avs {
// bytes of the packed form of a [rows, K] operand; tile = 128 for the activation side, 256 for the weight side
size_t gemm_packed_bytes(int rows, int K, int tile);
int gemm_pack(const float* x, int ld, int rows, int K, int tile, __nv_bfloat16* out, cudaStream_t st);
// C[M, N] = A . W^T + bias from packed operands (A packed with tile 128, W with tile 256); K % 32 == 0
int gemm_umma_nt(const __nv_bfloat16* a_packed, const __nv_bfloat16* w_packed, const float* bias, float* c, int ldc, int M,
                 int N, int K, int n_sms, cudaStream_t st);
// the same with K cut into `splits` slices (more CTAs for a small output): partial holds splits * M * N floats, ldc == N
int gemm_umma_nt_splitk(const __nv_bfloat16* a_packed, const __nv_bfloat16* w_packed, const float* bias, float* c, int ldc, int M,
                        int N, int K, int splits, float* partial, int n_sms, cudaStream_t st);
}  // namespace avs
