// fp32-grade GEMM on tcgen05 for the Bi-GRU input projections (K3):
//     C[M, N] = A[M, K] . W[N, K]^T + bias[N]         (A = activations, W = torch Linear / GRU weight)
// Every fp32 operand is split x = hi + lo in bf16 and each product is three MMAs
// (hi*hi + lo*hi + hi*lo, relative error ~2^-16), accumulated in fp32 in TMEM.
//
// Operands are pre-packed into the K-major no-swizzle UMMA layout in 8-element chunks:
//     P[kind = hi|lo][K/8 chunks][rows padded to the tile][8] bf16
// so a (chunk, 128- or 256-row) run is contiguous: tiles are loaded with plain cp.async.bulk and a
// K = 16 MMA reads two chunks LBO = rows*16 B apart.  Tile 128 x 256, BK = 32, 4-stage mbarrier ring,
// persistent CTAs (n fastest so an A tile is reused from L2 by its 6 column tiles), double-buffered
// TMEM accumulators, warp roles as in conv_umma.cu.
#include "common.cuh"
#include "gemm_umma.cuh"
#include "sgemm.cuh"

namespace avs {

constexpr int kGM = 128, kGN = 256, kGK = 32, kGStages = 4, kGThreads = 256;
constexpr int kAStage = 2 * (kGK / 8) * kGM * 16;  // hi + lo: 16 KB
constexpr int kBStage = 2 * (kGK / 8) * kGN * 16;  // 32 KB
constexpr int kStageBytes = kAStage + kBStage;
constexpr size_t kGemmSmem = static_cast<size_t>(kGStages) * kStageBytes + 256;

struct GemmParams {
  const __nv_bfloat16* a;  // packed [2][K8][Mp][8]
  const __nv_bfloat16* w;  // packed [2][K8][Np][8]
  const float* bias;
  float* c;
  int M, N, K8, Mp, Np, ldc, tiles_m, tiles_n;
  // split K (the detector's hidden layer: 16 output tiles, K = 13824): slice ks of an output tile covers the k-blocks
  // [ks * nk_split, (ks + 1) * nk_split) and writes its own dense [M, N] partial without bias; splits == 1: c + bias
  int splits, nk_split;
  float* partial;
};

__global__ void __launch_bounds__(kGThreads, 1)
gemm_umma_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(kGStages) * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kGStages;
  uint64_t* acc_full = empty + kGStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kGStages; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 4);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<512>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int n_out = p.tiles_m * p.tiles_n, n_tiles = n_out * p.splits, n_k = p.nk_split;

  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- producer
    uint32_t slot = 0, phase = 0;
    const size_t a_kind = static_cast<size_t>(p.K8) * p.Mp * 8, w_kind = static_cast<size_t>(p.K8) * p.Np * 8;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int ot = tile % n_out, kb0 = (tile / n_out) * n_k;
      const int m0 = (ot / p.tiles_n) * kGM, n0 = (ot % p.tiles_n) * kGN;
      for (int kb = kb0; kb < kb0 + n_k; ++kb) {
        mbar_wait(&empty[slot], phase ^ 1);
        mbar_expect_tx(&full[slot], kStageBytes);
        uint8_t* sa = smem + static_cast<size_t>(slot) * kStageBytes;
        uint8_t* sb = sa + kAStage;
#pragma unroll
        for (int kind = 0; kind < 2; ++kind)
#pragma unroll
          for (int c = 0; c < kGK / 8; ++c) {
            const size_t chunk = static_cast<size_t>(kb) * (kGK / 8) + c;
            bulk_g2s(sa + (kind * (kGK / 8) + c) * (kGM * 16), p.a + kind * a_kind + (chunk * p.Mp + m0) * 8, kGM * 16, &full[slot]);
            bulk_g2s(sb + (kind * (kGK / 8) + c) * (kGN * 16), p.w + kind * w_kind + (chunk * p.Np + n0) * 8, kGN * 16, &full[slot]);
          }
        if (++slot == kGStages) slot = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (converged warp, elected lane)
    const uint32_t idesc = umma_idesc_bf16(kGM, kGN);
    constexpr uint64_t kHi = static_cast<uint64_t>((128u >> 4) | (1u << 14)) << 32;           // SBO 128 B, version 1
    constexpr uint32_t kLboA = ((kGM * 16) >> 4) << 16, kLboB = ((kGN * 16) >> 4) << 16;
    const uint32_t base = (smem_u32(smem) & 0x3FFFFu) >> 4;
    uint32_t slot = 0, phase = 0, buf = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[buf], aphase ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + buf * kGN;
      for (int kb = 0; kb < n_k; ++kb) {
        mbar_wait(&full[slot], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = base + slot * (kStageBytes >> 4), sb = sa + (kAStage >> 4);
#pragma unroll
          for (int j = 0; j < kGK / 16; ++j) {
            const uint32_t a_hi = sa + (2 * j) * (kGM * 16 >> 4), a_lo = a_hi + (kGK / 8) * (kGM * 16 >> 4);
            const uint32_t b_hi = sb + (2 * j) * (kGN * 16 >> 4), b_lo = b_hi + (kGK / 8) * (kGN * 16 >> 4);
            umma_f16(d, kHi | kLboA | a_hi, kHi | kLboB | b_hi, idesc, (kb | j) != 0 ? 1u : 0u);
            umma_f16(d, kHi | kLboA | a_lo, kHi | kLboB | b_hi, idesc, 1u);
            umma_f16(d, kHi | kLboA | a_hi, kHi | kLboB | b_lo, idesc, 1u);
          }
          tc_commit(&empty[slot]);
          if (kb == n_k - 1) tc_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++slot == kGStages) slot = 0, phase ^= 1;
      }
      if (++buf == 2) buf = 0, aphase ^= 1;
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: + bias, store fp32 rows
    const int q = warp & 3;
    uint32_t buf = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int ot = tile % n_out, ks = tile / n_out;
      const int m0 = (ot / p.tiles_n) * kGM, n0 = (ot % p.tiles_n) * kGN;
      mbar_wait(&acc_full[buf], aphase);
      __syncwarp();  // tcgen05.ld below is .aligned
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool split = p.partial != nullptr;
      float* crow = split ? p.partial + (static_cast<size_t>(ks) * p.M + row) * p.N + n0 : p.c + static_cast<size_t>(row) * p.ldc + n0;
      const uint32_t d = tmem_base + buf * kGN + (static_cast<uint32_t>(q * 32) << 16);
      for (int cb = 0; cb < kGN; cb += 32) {
        uint32_t v[32];
        tmem_ld32(d + cb, v);
        tmem_ld_wait();
        if (row < p.M) {
#pragma unroll
          for (int c4 = 0; c4 < 32; c4 += 4) {
            const int n = n0 + cb + c4;
            if (n + 3 < p.N) {
              const float4 b4 = split ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(p.bias + n);
              *reinterpret_cast<float4*>(crow + cb + c4) =
                  make_float4(__uint_as_float(v[c4]) + b4.x, __uint_as_float(v[c4 + 1]) + b4.y,
                              __uint_as_float(v[c4 + 2]) + b4.z, __uint_as_float(v[c4 + 3]) + b4.w);
            } else {
              for (int e = 0; e < 4; ++e)
                if (n + e < p.N) crow[cb + c4 + e] = __uint_as_float(v[c4 + e]) + (split ? 0.f : p.bias[n + e]);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (++buf == 2) buf = 0, aphase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// fp32 [rows, K] row-major (leading dimension ld) -> packed hi/lo chunks [2][K/8][rows_p][8]; rows >= rows are zero.
// One thread per (chunk, row): consecutive threads write consecutive 16-byte positions.
__global__ void __launch_bounds__(256)
pack_split_kernel(const float* __restrict__ x, int ld, int rows, int rows_p, int K8, __nv_bfloat16* __restrict__ out) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= static_cast<long long>(K8) * rows_p) return;
  const int r = static_cast<int>(idx % rows_p), c = static_cast<int>(idx / rows_p);
  float v[8];
  if (r < rows) {
    const float4 a = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ld + c * 8);
    const float4 b = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ld + c * 8 + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
    hi[e] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
    lo[e] = pack_bf16x2(v[2 * e] - __bfloat162float(h0), v[2 * e + 1] - __bfloat162float(h1));
  }
  uint4* o = reinterpret_cast<uint4*>(out);
  o[idx] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  o[static_cast<long long>(K8) * rows_p + idx] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

size_t gemm_packed_bytes(int rows, int K, int tile) {
  return static_cast<size_t>(2) * (K / 8) * (static_cast<size_t>(cdiv(rows, tile)) * tile) * 16;
}

int gemm_pack(const float* x, int ld, int rows, int K, int tile, __nv_bfloat16* out, cudaStream_t st) {
  AVS_REQUIRE(K % kGK == 0 && ld % 4 == 0, "gemm_pack needs K % 32 == 0 and 16-byte aligned rows");
  const int rows_p = cdiv(rows, tile) * tile, K8 = K / 8;
  const long long total = static_cast<long long>(K8) * rows_p;
  pack_split_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(x, ld, rows, rows_p, K8, out);
  AVS_LAUNCHED();
  return AVS_OK;
}

int gemm_umma_nt(const __nv_bfloat16* a_packed, const __nv_bfloat16* w_packed, const float* bias, float* c, int ldc, int M,
                 int N, int K, int n_sms, cudaStream_t st) {
  return gemm_umma_nt_splitk(a_packed, w_packed, bias, c, ldc, M, N, K, 1, nullptr, n_sms, st);
}

// splits > 1: `partial` holds splits * M * N floats, c is dense (ldc == N); the slices are summed in slice order by
// splitk_reduce_kernel (deterministic: a row's result does not depend on M or on the row's place in its tile)
int gemm_umma_nt_splitk(const __nv_bfloat16* a_packed, const __nv_bfloat16* w_packed, const float* bias, float* c, int ldc, int M,
                        int N, int K, int splits, float* partial, int n_sms, cudaStream_t st) {
  AVS_REQUIRE(K % kGK == 0 && ldc % 4 == 0 && N % 4 == 0, "gemm_umma_nt shape");
  AVS_REQUIRE(splits >= 1 && (K / kGK) % splits == 0 && (splits == 1 || (partial && ldc == N)), "gemm_umma_nt split-K configuration");
  // per-device attribute: set on every call (cheap), a process may drive several GPUs
  AVS_CUDA(cudaFuncSetAttribute(gemm_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem)));
  GemmParams p;
  p.a = a_packed; p.w = w_packed; p.bias = bias; p.c = c;
  p.M = M; p.N = N; p.K8 = K / 8; p.ldc = ldc;
  p.tiles_m = cdiv(M, kGM); p.tiles_n = cdiv(N, kGN);
  p.Mp = p.tiles_m * kGM; p.Np = p.tiles_n * kGN;
  p.splits = splits; p.nk_split = K / kGK / splits; p.partial = splits > 1 ? partial : nullptr;
  const int grid = std::min(p.tiles_m * p.tiles_n * splits, n_sms);
  gemm_umma_kernel<<<grid, kGThreads, kGemmSmem, st>>>(p);
  AVS_LAUNCHED();
  if (splits > 1) return splitk_reduce(partial, bias, c, M, N, splits, st);
  return AVS_OK;
}

}  // namespace avs

// C-ABI wrapper (tests, and the DFT-as-GEMM comparison of tools/k1_gemm_vs_fft.py): packs both operands and
// runs the split GEMM.  workspace >= avs_gemm_split_workspace_bytes(M, N, K).
extern "C" size_t avs_gemm_split_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return avs::align_up(avs::gemm_packed_bytes(M, K, 128), 256) + avs::align_up(avs::gemm_packed_bytes(N, K, 256), 256);
}

extern "C" int avs_gemm_split(const float* a, const float* w, const float* bias, float* c, int M, int N, int K,
                              void* workspace, size_t workspace_bytes, void* stream) {
  using namespace avs;
  AVS_REQUIRE(a && w && bias && c && workspace, "null argument");
  AVS_REQUIRE(K % 32 == 0 && N % 4 == 0, "avs_gemm_split needs K % 32 == 0 and N % 4 == 0");
  if (workspace_bytes < avs_gemm_split_workspace_bytes(M, N, K)) {
    set_error("gemm_split workspace too small");
    return AVS_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, n_sms = 148;
  AVS_CUDA(cudaGetDevice(&dev));
  AVS_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  __nv_bfloat16* ap = static_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(workspace) + align_up(gemm_packed_bytes(M, K, 128), 256));
  int rc;
  if ((rc = gemm_pack(a, K, M, K, 128, ap, st))) return rc;
  if ((rc = gemm_pack(w, K, N, K, 256, wp, st))) return rc;
  return gemm_umma_nt(ap, wp, bias, c, N, M, N, K, n_sms, st);
}
