// K3 recurrence on the tensor cores (hidden = 256): the per-step mat-vec  gh[rows, clips] = W_hh[rows, :] . h[clips, :]^T
// of one cluster CTA (96 gate rows of 32 hidden units, 16 clips) is 32 tcgen05.mma per step and clip group.
//
//   * W_hh slice resident in shared memory for the whole sequence, split hi/lo in bf16, in the K-major
//     no-swizzle UMMA layout [k chunk][128 rows][8] (rows 96..127 are zero padding): the A operand.
//   * h(t-1) of a group of 16 clips lives in every CTA of the cluster as the B operand, also split hi/lo:
//     [k chunk][hi clips 0..15 | lo clips 0..15][8], double buffered.  One MMA of width 32 computes
//     W_hi.h_hi and W_hi.h_lo, a second one adds W_lo.[h_hi | h_lo] (fp32-grade, like gemm_umma.cu).
//   * a cluster serves TWO groups of 16 clips, software-pipelined: a dedicated warp issues the MMAs of a group as soon as
//     its h(t-1) has arrived from all eight CTAs, while the eight worker warps read the other group's accumulator
//     (tcgen05.ld, lane = gate row), exchange through shared memory so that thread (unit, clip pair) holds r, z, n,
//     evaluate the gates in fp32 and send the CTA's 32 new hidden values to all 8 CTAs' B operands as bulk DSMEM copies.
//     The step of one group (MMAs 1400 cycles, TMEM read 500, gates + staging 1600, exchange + wait 700 and the barrier
//     latencies between them) is a dependent chain of ~8000 cycles; with two groups in flight the tensor core and the
//     exchange of one group run under the gate arithmetic of the other, and 256 clips are ONE wave of 16 clusters x 8
//     CTAs x 2 directions instead of two.
// Cluster of 8 CTAs per (32 clips, direction), as in gru.cu's CUDA-core kernel (which remains the fp32 path).
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"
#include "gru_umma.cuh"

namespace avs {

constexpr int kH = 256, kClu = 8, kUnits = 32, kRows = 128, kChunks = kH / 8;
constexpr int kABytes = kChunks * kRows * 16;             // one kind (hi or lo): 64 KB
constexpr int kGrpClips = 16, kGroups = 2;                // clips per group, groups per cluster
constexpr int kBBytes = kChunks * 2 * kGrpClips * 16;     // one h buffer (hi + lo) of one group: 16 KB
constexpr int kXsPitch = kGrpClips + 1;
constexpr int kStageBytes = 4 * 2 * kGrpClips * 16;       // this CTA's 32 units (4 chunks) of h, hi + lo: 2 KB
#ifndef AVS_VAR_GRU_WORKERS
#define AVS_VAR_GRU_WORKERS 16
#endif
constexpr int kWorkers = AVS_VAR_GRU_WORKERS;             // worker warps: 16 = one clip of the group per warp (8: two)
constexpr int kCpt = kGrpClips / kWorkers;                // clips per thread and group (thread = (unit, warp's clips))
constexpr int kGruThreads = (kWorkers + 1) * 32;          // worker warps + the MMA-issuing warp
constexpr size_t kGruSmem = 2ull * kABytes + kGroups * 2ull * kBBytes + kGroups * kRows * kXsPitch * 4 +
                            kGroups * 2 * kStageBytes + 128;

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
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkers * 32) : "memory"); }  // the worker warps

// wp: packed W_hh [2 dirs][8 ranks][2 kinds][32 chunks][128 rows][8] bf16 (gru_pack_whh)
//
// Step protocol (no cluster barrier inside the loop): h(t) travels between CTAs as bulk async copies that
// complete on the RECEIVER's mbarrier bar_h[group][buffer] (8 x 2 KB per step and group), so the tensor core only ever
// reads operand bytes written through the async proxy.  Buffer reuse is safe by data dependence: a CTA can start
// step s+1 of a group only after every peer delivered h(s), which each peer sends after its own step-s MMAs finished.
__global__ void __launch_bounds__(kGruThreads, 1)
gru_cluster_umma_kernel(const float* __restrict__ xp, const __nv_bfloat16* __restrict__ wp, const float* __restrict__ b_hh,
                        float* __restrict__ out, int B, int T) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_a = smem;                                               // [hi|lo][chunk][row][16 B]
  uint8_t* s_b = s_a + 2 * kABytes;                                  // [group][2 buffers][chunk][hi|lo][clip][16 B]
  float* s_x = reinterpret_cast<float*>(s_b + kGroups * 2 * kBBytes);   // [group][128 rows][17]
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_x + kGroups * kRows * kXsPitch);  // [group][2][this CTA's 4 chunks: chunk][hi|lo][clip][16 B]
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_stage + kGroups * 2 * kStageBytes);
  uint64_t* bar_mma = bar_w + 1;                                     // [group]
  uint64_t* bar_h = bar_mma + kGroups;                               // [group][2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_h + kGroups * 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cta_rank();
  const int dir = blockIdx.y, cluster = blockIdx.x / kClu;
  const int J = rank * kUnits + lane;
  const int clip0 = cluster * kGroups * kGrpClips;                   // first clip of the cluster
  // groups with at least one clip (uniform over the cluster): the second group of the last cluster may be empty
  const int n_groups = (B - clip0 > kGrpClips) ? kGroups : 1;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int g = 0; g < kGroups; ++g) {
      mbar_init(&bar_mma[g], 1);
      mbar_init(&bar_h[g * 2 + 0], 1);
      mbar_init(&bar_h[g * 2 + 1], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<64>(s_tmem);  // a warp that has not diverged: tcgen05.alloc is .sync.aligned; 2 groups x 32 columns
  for (int i = tid; i < kGroups * 2 * kBBytes / 16; i += kGruThreads) reinterpret_cast<uint4*>(s_b)[i] = make_uint4(0, 0, 0, 0);  // h(-1) = 0
  fence_proxy_async();  // generic-proxy zero fill, async-proxy (tcgen05.mma) reader
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *s_tmem;
  if (tid == 0) {  // resident weights: 128 KB in 8 bulk copies; and expect h(0) of every group in its buffer 1
    const uint8_t* src = reinterpret_cast<const uint8_t*>(wp) + (static_cast<size_t>(dir) * kClu + rank) * 2 * kABytes;
    mbar_expect_tx(bar_w, 2 * kABytes);
    for (int i = 0; i < 8; ++i) bulk_g2s(s_a + i * (kABytes / 4), src + static_cast<size_t>(i) * (kABytes / 4), kABytes / 4, bar_w);
    for (int g = 0; g < n_groups; ++g) mbar_expect_tx(&bar_h[g * 2 + 1], kBBytes);
  }
  mbar_wait(bar_w, 0);
  // all CTAs of the cluster have initialised their barriers and buffers before anyone sends
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");

  if (warp == kWorkers) {
    // ============================================================ MMA issuer
    // In a cluster launch the shared-window address of a CTA carries its cluster rank above bit 24 (rank 1: 0x01000400).
    // The descriptor's start-address field is 14 bits of (address >> 4): mask, or the rank lands in the LBO field.
    const uint32_t a_lo32 = (smem_u32(s_a) & 0x3FFFFu) >> 4, b_lo32 = (smem_u32(s_b) & 0x3FFFFu) >> 4;
    constexpr uint64_t kHi = static_cast<uint64_t>((128u >> 4) | (1u << 14)) << 32;  // SBO 128 B, descriptor version 1
    constexpr uint32_t kLboA = ((kRows * 16) >> 4) << 16, kLboB = ((2 * kGrpClips * 16) >> 4) << 16;
    const uint32_t idesc_w = umma_idesc_bf16(128, 2 * kGrpClips);
    uint32_t h_phase = 0;  // bit (g * 2 + buffer)
    for (int s = 0; s < T; ++s) {
      const int cur = s & 1;
      for (int g = 0; g < n_groups; ++g) {
        if (s > 0) {  // h(t-1) of this group from all 8 CTAs has landed in buffer `cur`
          mbar_wait(&bar_h[g * 2 + cur], (h_phase >> (g * 2 + cur)) & 1u);
          h_phase ^= 1u << (g * 2 + cur);
        }
        tc_fence_after();
        if (elect_one()) {  // converged warp, one elected lane issues
          if (s + 2 < T) mbar_expect_tx(&bar_h[g * 2 + cur], kBBytes);  // next tenant of this buffer: h(t+1), sent during step s+1
          const uint32_t bb = b_lo32 + (g * 2 + cur) * (kBBytes >> 4);
          const uint32_t d = tmem_d + g * 2 * kGrpClips;
#pragma unroll
          for (int j = 0; j < kChunks / 2; ++j) {
            const uint32_t a_hi = a_lo32 + (2 * j) * (kRows * 16 >> 4), a_lo = a_hi + (kABytes >> 4);
            const uint32_t bj = bb + (2 * j) * (2 * kGrpClips * 16 >> 4);
            umma_f16(d, kHi | kLboA | a_hi, kHi | kLboB | bj, idesc_w, j != 0 ? 1u : 0u);
            umma_f16(d, kHi | kLboA | a_lo, kHi | kLboB | bj, idesc_w, 1u);  // also adds the tiny W_lo.h_lo term
          }
          tc_commit(&bar_mma[g]);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================================================ workers: TMEM read-out, gates, exchange
    const float br = b_hh[dir * 3 * kH + J], bz = b_hh[dir * 3 * kH + kH + J], bn = b_hh[dir * 3 * kH + 2 * kH + J];
    const int c0 = kCpt * warp;                 // this thread's clips inside a group: c0 .. c0 + kCpt - 1
    float h[kGroups][kCpt];
#pragma unroll
    for (int g = 0; g < kGroups; ++g)
#pragma unroll
      for (int k = 0; k < kCpt; ++k) h[g][k] = 0.f;
    for (int s = 0; s < T; ++s) {
      const int t = dir ? T - 1 - s : s;
      const int cur = s & 1;
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        if (g >= n_groups) break;
        const int b0 = clip0 + g * kGrpClips + c0;
        const int n_valid = min(kGrpClips, B - clip0 - g * kGrpClips);
        // input-projection terms: their latency hides behind the wait for the MMAs
        float gi[kCpt][3];
#pragma unroll
        for (int k = 0; k < kCpt; ++k) {
          gi[k][0] = gi[k][1] = gi[k][2] = 0.f;
          if (b0 + k < B) {
            const float* gp = xp + (static_cast<size_t>(b0 + k) * T + t) * 6 * kH + dir * 3 * kH + J;
            gi[k][0] = gp[0]; gi[k][1] = gp[kH]; gi[k][2] = gp[2 * kH];
          }
        }
        float* sx = s_x + g * kRows * kXsPitch;
        if (warp < 4) {  // lane of TMEM = gate row (gate * 32 + unit); rows >= 96 are padding
          mbar_wait(&bar_mma[g], s & 1);
          __syncwarp();
          tc_fence_after();
          float* xr = sx + (warp * 32 + lane) * kXsPitch;
          uint32_t v[32];
          tmem_ld32(tmem_d + g * 2 * kGrpClips + (static_cast<uint32_t>(warp * 32) << 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) xr[c] = __uint_as_float(v[c]) + __uint_as_float(v[16 + c]);
          tc_fence_before();
        }
        workers_sync();
        // ---- gates: thread = (unit, the warp's kCpt clips)
#pragma unroll
        for (int k = 0; k < kCpt; ++k) {
          if (c0 + k < n_valid) {
            const int c = c0 + k;
            const float r = 1.f / (1.f + expf(-(gi[k][0] + sx[(0 * kUnits + lane) * kXsPitch + c] + br)));
            const float z = 1.f / (1.f + expf(-(gi[k][1] + sx[(1 * kUnits + lane) * kXsPitch + c] + bz)));
            const float n = tanhf(gi[k][2] + r * (sx[(2 * kUnits + lane) * kXsPitch + c] + bn));
            h[g][k] = (1.f - z) * n + z * h[g][k];
            out[(static_cast<size_t>(b0 + k) * T + t) * 2 * kH + dir * kH + J] = h[g][k];
          }
        }
        if (s + 1 < T) {
          // this CTA's 32 units of h(t), split hi/lo, staged in operand layout [chunk (4)][hi|lo][clip][8] ...
          uint8_t* stage = s_stage + (g * 2 + cur) * kStageBytes;
          __nv_bfloat16* st = reinterpret_cast<__nv_bfloat16*>(stage) + (lane >> 3) * (2 * kGrpClips * 8) + (lane & 7);
#pragma unroll
          for (int k = 0; k < kCpt; ++k) {
            const __nv_bfloat16 hh = __float2bfloat16_rn(h[g][k]);
            st[(c0 + k) * 8] = hh;
            st[kGrpClips * 8 + (c0 + k) * 8] = __float2bfloat16_rn(h[g][k] - __bfloat162float(hh));
          }
          fence_proxy_async();  // staged with generic stores, read by the bulk-copy engine
          workers_sync();
          // ... and pushed to chunks [4 rank, 4 rank + 4) of every CTA's next buffer of this group: one bulk copy per CTA
          if (tid < kClu) {  // (lanes 0..7 of worker warp 0)
            const uint32_t dst_off = static_cast<uint32_t>((g * 2 + (cur ^ 1)) * kBBytes + rank * kStageBytes);
            bulk_s2c(map_to_rank(smem_u32(s_b) + dst_off, tid), stage, kStageBytes, map_to_rank(smem_u32(&bar_h[g * 2 + (cur ^ 1)]), tid));
          }
        }
      }
    }
  }
  // nobody may exit while a peer could still be sending to it or reading its staging buffers
  __syncwarp();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<64>(tmem_d);
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

int gru_recurrence_umma(const float* xp, const __nv_bfloat16* w_packed, const float* b_hh, float* out, int B, int T,
                        cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kClu * cdiv(B, kGroups * kGrpClips), 2, 1);
  cfg.blockDim = dim3(kGruThreads, 1, 1);
  cfg.dynamicSmemBytes = kGruSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClu; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  AVS_CUDA(cudaFuncSetAttribute(gru_cluster_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGruSmem)));
  AVS_CUDA(cudaLaunchKernelEx(&cfg, gru_cluster_umma_kernel, xp, w_packed, b_hh, out, B, T));
  AVS_LAUNCHED();
  return AVS_OK;
}

#ifdef AVS_EXPERIMENTS
// tools: how many 8-CTA clusters of the recurrence kernel the device can hold at once
int gru_max_active_clusters() {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kClu * 64, 2, 1);
  cfg.blockDim = dim3(kGruThreads, 1, 1);
  cfg.dynamicSmemBytes = kGruSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClu; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaFuncSetAttribute(gru_cluster_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGruSmem));
  int n = -1;
  if (cudaOccupancyMaxActiveClusters(&n, gru_cluster_umma_kernel, &cfg) != cudaSuccess) n = -1;
  return n;
}
#endif

}  // namespace avs
#ifdef AVS_EXPERIMENTS
extern "C" __attribute__((visibility("default"))) int avs_gru_max_active_clusters(void) { return avs::gru_max_active_clusters(); }
#endif
