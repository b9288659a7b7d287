// tcgen05 recurrence kernel of the Bi-GRU head (gru_umma.cu), hidden = 256 only
#pragma once
#include "common.cuh"
namespace avs {
size_t gru_whh_packed_bytes();
int gru_pack_whh(const float* w_hh /* [2][3H][H] */, __nv_bfloat16* out, cudaStream_t st);
int gru_recurrence_umma(const float* xp, const __nv_bfloat16* w_packed, const float* b_hh, float* out, int B, int T,
                        cudaStream_t st);
}  // namespace avs
