// K3 recurrence on the tensor cores (hidden = 256): the per-step mat-vec  gh[rows, clips] = W_hh[rows, :] . h[clips, :]^T
// of one cluster CTA (96 gate rows of 32 hidden units, 16 or 32 clips) is 32 tcgen05.mma per step.
//
//   * W_hh slice resident in shared memory for the whole sequence, split hi/lo in bf16, in the K-major
//     no-swizzle UMMA layout [k chunk][128 rows][8] (rows 96..127 are zero padding): the A operand.
//   * h(t-1) of the 16 clips lives in every CTA of the cluster as the B operand, also split hi/lo:
//     [k chunk][hi clips 0..15 | lo clips 0..15][8], double buffered.  One MMA of width 32 computes
//     W_hi.h_hi and W_hi.h_lo, one of width 16 adds W_lo.h_hi (fp32-grade, like gemm_umma.cu).
//   * per step: cluster barrier -> 32 MMAs + commit -> 4 warps read the accumulator (lane = gate row),
//     exchange through shared memory so that thread (unit, clip pair) holds r, z, n -> gates in fp32 ->
//     the CTA's 32 new hidden values go to all 8 CTAs' B operands as 16-byte DSMEM stores.
// Cluster of 8 CTAs per (16 clips, direction), as in gru.cu's CUDA-core kernel (which remains the fp32 path).
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"
#include "gru_umma.cuh"

namespace avs {

constexpr int kH = 256, kClu = 8, kUnits = 32, kRows = 128, kChunks = kH / 8;
constexpr int kABytes = kChunks * kRows * 16;             // one kind (hi or lo): 64 KB
// CLIPS = clips per cluster (16 is what runs; the kernel also works with 32)
template <int CLIPS>
struct GruCfg {
  static constexpr int kBBytes = kChunks * 2 * CLIPS * 16;      // one h buffer (hi + lo): 16 / 32 KB
  static constexpr int kXsPitch = CLIPS + 1;
  static constexpr int kStageBytes = 4 * 2 * CLIPS * 16;        // this CTA's 32 units (4 chunks) of h, hi + lo: 2 / 4 KB
  static constexpr int kCpt = CLIPS / 8;                        // clips per thread (thread = (unit, warp's clips))
  static constexpr int kTmemCols = 2 * CLIPS;                   // D = [W.h_hi | W.h_lo]
  static constexpr size_t kSmem = 2ull * kABytes + 2ull * kBBytes + kRows * kXsPitch * 4 + 2 * kStageBytes + 64;
};

__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  return remote;
}
// bulk copy local shared memory -> shared memory of another CTA of the cluster; completes on THAT CTA's mbarrier
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster_addr, const void* src_smem, uint32_t bytes, uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
               "r"(smem_u32(src_smem)), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}

// wp: packed W_hh [2 dirs][8 ranks][2 kinds][32 chunks][128 rows][8] bf16 (gru_pack_whh)
//
// Step protocol (no cluster barrier inside the loop): h(t) travels between CTAs as bulk async copies that
// complete on the RECEIVER's mbarrier bar_h[buffer] (8 x 2 KB per step), so the tensor core only ever reads
// operand bytes written through the async proxy.  Buffer reuse is safe by data dependence: a CTA can start
// step s+1 only after every peer delivered h(s), which each peer sends after its own step-s MMAs finished.
template <int CLIPS>
__global__ void __launch_bounds__(256, 1)
gru_cluster_umma_kernel(const float* __restrict__ xp, const __nv_bfloat16* __restrict__ wp, const float* __restrict__ b_hh,
                        float* __restrict__ out, int B, int T) {
  using C = GruCfg<CLIPS>;
  constexpr int kBBytes = C::kBBytes, kXsPitch = C::kXsPitch, kStageBytes = C::kStageBytes, kCpt = C::kCpt;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_a = smem;                                               // [hi|lo][chunk][row][16 B]
  uint8_t* s_b = s_a + 2 * kABytes;                                  // [2 buffers][chunk][hi|lo][clip][16 B]
  float* s_x = reinterpret_cast<float*>(s_b + 2 * kBBytes);          // [128 rows][CLIPS + 1]
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_x + kRows * kXsPitch);  // [2][this CTA's 4 chunks: chunk][hi|lo][clip][16 B]
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_stage + 2 * kStageBytes);
  uint64_t* bar_mma = bar_w + 1;
  uint64_t* bar_h = bar_mma + 1;                                     // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_h + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cta_rank();
  const int dir = blockIdx.y, group = blockIdx.x / kClu;
  const int J = rank * kUnits + lane;
  const int n_valid = min(CLIPS, B - group * CLIPS);
  const int c0 = kCpt * warp;                 // this thread's clips inside the group: c0 .. c0 + kCpt - 1
  const int b0 = group * CLIPS + c0;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    mbar_init(&bar_h[0], 1);
    mbar_init(&bar_h[1], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(s_tmem);  // a warp that has not diverged: tcgen05.alloc is .sync.aligned
  for (int i = tid; i < kBBytes / 16; i += 256) reinterpret_cast<uint4*>(s_b)[i] = make_uint4(0, 0, 0, 0);  // h(-1) = 0 in buffer 0
  fence_proxy_async();  // generic-proxy zero fill, async-proxy (tcgen05.mma) reader
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *s_tmem;
  if (tid == 0) {  // resident weights: 128 KB in 8 bulk copies; and expect h(0) in buffer 1
    const uint8_t* src = reinterpret_cast<const uint8_t*>(wp) + (static_cast<size_t>(dir) * kClu + rank) * 2 * kABytes;
    mbar_expect_tx(bar_w, 2 * kABytes);
    for (int i = 0; i < 8; ++i) bulk_g2s(s_a + i * (kABytes / 4), src + static_cast<size_t>(i) * (kABytes / 4), kABytes / 4, bar_w);
    mbar_expect_tx(&bar_h[1], kBBytes);
  }
  const float br = b_hh[dir * 3 * kH + J], bz = b_hh[dir * 3 * kH + kH + J], bn = b_hh[dir * 3 * kH + 2 * kH + J];
  float h[kCpt];
#pragma unroll
  for (int k = 0; k < kCpt; ++k) h[k] = 0.f;
  mbar_wait(bar_w, 0);
  // all CTAs of the cluster have initialised their barriers and buffers before anyone sends
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");

  // In a cluster launch the shared-window address of a CTA carries its cluster rank above bit 24 (rank 1: 0x01000400).
  // The descriptor's start-address field is 14 bits of (address >> 4): mask, or the rank lands in the LBO field.
  const uint32_t a_lo32 = (smem_u32(s_a) & 0x3FFFFu) >> 4, b_lo32 = (smem_u32(s_b) & 0x3FFFFu) >> 4;
  constexpr uint64_t kHi = static_cast<uint64_t>((128u >> 4) | (1u << 14)) << 32;  // SBO 128 B, descriptor version 1
  constexpr uint32_t kLboA = ((kRows * 16) >> 4) << 16, kLboB = ((2 * CLIPS * 16) >> 4) << 16;
  const uint32_t idesc_w = umma_idesc_bf16(128, 2 * CLIPS);
  uint32_t h_phase[2] = {0, 0};

  for (int s = 0; s < T; ++s) {
    const int t = dir ? T - 1 - s : s;
    const int cur = s & 1;
    // ---- mat-vec on the tensor core: D[row, 0:CLIPS] = (W_hi + W_lo).h_hi, D[row, CLIPS:2 CLIPS] = (W_hi + W_lo).h_lo
    if (warp == 0) {  // converged warp, one elected lane issues: a tcgen05.mma inside a divergent branch costs ~49 instead of 40 cycles
      if (s > 0) {  // h(t-1) from all 8 CTAs has landed in buffer `cur`
        mbar_wait(&bar_h[cur], h_phase[cur]);
        h_phase[cur] ^= 1;
      }
      tc_fence_after();
      if (elect_one()) {
        if (s + 2 < T) mbar_expect_tx(&bar_h[cur], kBBytes);  // next tenant of this buffer: h(t+1), sent during step s+1
        const uint32_t bb = b_lo32 + cur * (kBBytes >> 4);
#pragma unroll
        for (int j = 0; j < kChunks / 2; ++j) {
          const uint32_t a_hi = a_lo32 + (2 * j) * (kRows * 16 >> 4), a_lo = a_hi + (kABytes >> 4);
          const uint32_t bj = bb + (2 * j) * (2 * CLIPS * 16 >> 4);
          umma_f16(tmem_d, kHi | kLboA | a_hi, kHi | kLboB | bj, idesc_w, j != 0 ? 1u : 0u);
          umma_f16(tmem_d, kHi | kLboA | a_lo, kHi | kLboB | bj, idesc_w, 1u);  // also adds the tiny W_lo.h_lo term
        }
        tc_commit(bar_mma);
      }
      __syncwarp();
    }
    // input-projection terms: their latency hides behind the MMAs
    float gi[kCpt][3];
#pragma unroll
    for (int k = 0; k < kCpt; ++k) {
      gi[k][0] = gi[k][1] = gi[k][2] = 0.f;
      if (b0 + k < B) {
        const float* g = xp + (static_cast<size_t>(b0 + k) * T + t) * 6 * kH + dir * 3 * kH + J;
        gi[k][0] = g[0]; gi[k][1] = g[kH]; gi[k][2] = g[2 * kH];
      }
    }
    if (warp < 4) {  // lane of TMEM = gate row (gate * 32 + unit); rows >= 96 are padding
      mbar_wait(bar_mma, s & 1);
      __syncwarp();  // lane 0 issued the MMAs and arrives late: tcgen05.ld is .aligned and needs the warp converged
      tc_fence_after();
      float* xr = s_x + (warp * 32 + lane) * kXsPitch;
      const uint32_t trow = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
      if (CLIPS == 16) {
        uint32_t v[32];
        tmem_ld32(trow, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) xr[c] = __uint_as_float(v[c]) + __uint_as_float(v[16 + c]);
      } else {
        uint32_t v0[32], v1[32];
        tmem_ld32(trow, v0);
        tmem_ld32(trow + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) xr[c] = __uint_as_float(v0[c]) + __uint_as_float(v1[c]);
      }
      tc_fence_before();
    }
    __syncthreads();
    // ---- gates: thread = (unit, the warp's kCpt clips)
#pragma unroll
    for (int k = 0; k < kCpt; ++k) {
      if (c0 + k < n_valid) {
        const int c = c0 + k;
        const float r = 1.f / (1.f + expf(-(gi[k][0] + s_x[(0 * kUnits + lane) * kXsPitch + c] + br)));
        const float z = 1.f / (1.f + expf(-(gi[k][1] + s_x[(1 * kUnits + lane) * kXsPitch + c] + bz)));
        const float n = tanhf(gi[k][2] + r * (s_x[(2 * kUnits + lane) * kXsPitch + c] + bn));
        h[k] = (1.f - z) * n + z * h[k];
        out[(static_cast<size_t>(b0 + k) * T + t) * 2 * kH + dir * kH + J] = h[k];
      }
    }
    if (s + 1 < T) {
      // this CTA's 32 units of h(t), split hi/lo, staged in operand layout [chunk (4)][hi|lo][clip][8] ...
      uint8_t* stage = s_stage + cur * kStageBytes;
      __nv_bfloat16* st = reinterpret_cast<__nv_bfloat16*>(stage) + (lane >> 3) * (2 * CLIPS * 8) + (lane & 7);
#pragma unroll
      for (int k = 0; k < kCpt; ++k) {
        const __nv_bfloat16 hh = __float2bfloat16_rn(h[k]);
        st[(c0 + k) * 8] = hh;
        st[CLIPS * 8 + (c0 + k) * 8] = __float2bfloat16_rn(h[k] - __bfloat162float(hh));
      }
      fence_proxy_async();  // staged with generic stores, read by the bulk-copy engine
      __syncthreads();
      // ... and pushed to chunks [4 rank, 4 rank + 4) of every CTA's next buffer: one bulk copy per CTA
      if (tid < kClu) {
        const uint32_t dst_off = static_cast<uint32_t>((cur ^ 1) * kBBytes + rank * kStageBytes);
        bulk_s2c(map_to_rank(smem_u32(s_b) + dst_off, tid), stage, kStageBytes, map_to_rank(smem_u32(&bar_h[cur ^ 1]), tid));
      }
    }
  }
  // nobody may exit while a peer could still be sending to it or reading its staging buffers
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_d);
  }
}

// w_hh [2][3H][H] f32 (reference layout) -> [2 dirs][8 ranks][hi|lo][32 chunks][128 rows][8] bf16
__global__ void __launch_bounds__(256)
gru_pack_whh_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * 256 + threadIdx.x;  // one thread per (dir, rank, chunk, row)
  if (idx >= 2 * kClu * kChunks * kRows) return;
  const int row = idx % kRows, chunk = (idx / kRows) % kChunks, rank = (idx / (kRows * kChunks)) % kClu,
            dir = idx / (kRows * kChunks * kClu);
  uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
  if (row < 3 * kUnits) {
    const int gate = row / kUnits, unit = rank * kUnits + row % kUnits;
    const float* src = w + (static_cast<size_t>(dir) * 3 * kH + gate * kH + unit) * kH + chunk * 8;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = src[2 * e], x1 = src[2 * e + 1];
      const __nv_bfloat16 a = __float2bfloat16_rn(x0), b = __float2bfloat16_rn(x1);
      hi[e] = static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
      lo[e] = pack_bf16x2(x0 - __bfloat162float(a), x1 - __bfloat162float(b));
    }
  }
  uint4* o = reinterpret_cast<uint4*>(out) + (static_cast<size_t>(dir) * kClu + rank) * 2 * kChunks * kRows;
  o[chunk * kRows + row] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  o[kChunks * kRows + chunk * kRows + row] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

size_t gru_whh_packed_bytes() { return static_cast<size_t>(2) * kClu * 2 * kABytes; }

int gru_pack_whh(const float* w_hh, __nv_bfloat16* out, cudaStream_t st) {
  gru_pack_whh_kernel<<<cdiv(2 * kClu * kChunks * kRows, 256), 256, 0, st>>>(w_hh, out);
  AVS_LAUNCHED();
  return AVS_OK;
}

template <int CLIPS>
static int launch_recurrence(const float* xp, const __nv_bfloat16* w_packed, const float* b_hh, float* out, int B, int T,
                             cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kClu * cdiv(B, CLIPS), 2, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = GruCfg<CLIPS>::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClu; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  AVS_CUDA(cudaFuncSetAttribute(gru_cluster_umma_kernel<CLIPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(GruCfg<CLIPS>::kSmem)));
  AVS_CUDA(cudaLaunchKernelEx(&cfg, gru_cluster_umma_kernel<CLIPS>, xp, w_packed, b_hh, out, B, T));
  AVS_LAUNCHED();
  return AVS_OK;
}

int gru_recurrence_umma(const float* xp, const __nv_bfloat16* w_packed, const float* b_hh, float* out, int B, int T,
                        cudaStream_t st) {
  // 16 clips per cluster.  32 (half the clusters: one wave instead of two at 256 clips) was measured too: its step takes
  // 7.6 us against 4.2, so it only breaks even — the step is gate- and exchange-bound, not MMA-bound.
  return launch_recurrence<16>(xp, w_packed, b_hh, out, B, T, st);
}

}  // namespace avs
