// Micro-benchmark: sustained cost of one tcgen05.mma (M = 128, K = 16, bf16 -> f32, cta_group::1) as a function
// of N, of the shared-memory operand layout and of how many accumulators the stream rotates over.
// One CTA per SM on every SM, one issuing thread, operands resident in shared memory (no loads in the loop).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I<csrc> -o umma_rate tools/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"

using namespace avs;

struct Variant {
  int N, layout, n_acc, reps;  // layout 0: K-major no-swizzle (8x16B core matrices), 1: K-major SWIZZLE_128B
  int a_tiles, b_tiles;        // distinct operand tiles the stream rotates over
  int table;                   // 1: descriptors built per MMA from a shared-memory schedule table, as in conv_umma.cu (2 tiles per entry)
  int delay, commit;
  int poll;                    // 0: nobody else on the SM; 1: 8 more warps spin on an mbarrier with all lanes; 2: lane 0 only           // every 24 MMAs: spin `delay` cycles on the issuing thread / issue a tcgen05.commit
};

__global__ void __launch_bounds__(384, 1) rate_kernel(Variant v, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint4 s_tab[24];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x;
  for (int i = tid; i < 200 * 1024 / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ (i * 2654435761u & 0x00ff00ffu);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); mbar_fence_init(); }
  if (tid < 24) s_tab[tid] = make_uint4((tid % 8) * 8 + ((2048u >> 4) << 16), (tid % 4) * ((v.N * 32) >> 4) + (((v.N * 16u) >> 4) << 16), 0, tid == 99 ? 32 : 0);
  if (tid < 32) tmem_alloc<512>(&s_tmem);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s_tmem;
  if (tid < 32) {  // converged warp, one elected lane issues (no per-MMA uniformity loop in the SASS)
    const uint32_t idesc = umma_idesc_bf16(128, v.N);
    const uint32_t base = (smem_u32(smem) & 0x3FFFFu) >> 4;
    uint64_t hi;
    uint32_t a_step, b_step, a_lbo = 0, b_lbo = 0;
    if (v.layout == 0) {
      hi = static_cast<uint64_t>((128u >> 4) | (1u << 14)) << 32;
      a_lbo = (2048u >> 4) << 16;                 // [K half][128 rows][16 B]
      b_lbo = ((v.N * 16u) >> 4) << 16;
      a_step = 4096 >> 4; b_step = (v.N * 32) >> 4;
    } else {
      hi = (static_cast<uint64_t>((1024u >> 4) | (1u << 14)) << 32) | (static_cast<uint64_t>(2) << 61);
      a_lbo = 1u << 16; b_lbo = 1u << 16;
      a_step = 32 >> 4; b_step = 32 >> 4;          // next K = 16 slice inside the 128-byte swizzle atom
    }
    const uint32_t a0 = base, b0 = base + (96 * 1024 >> 4);
    const uint32_t acc_cols = 512 / v.n_acc >= v.N ? 512 / v.n_acc : v.N;
    // descriptors of 8 consecutive MMAs precomputed: the timed loop is 8 back-to-back tcgen05.mma per iteration
    uint64_t da[8], db[8];
    uint32_t dd[8];
    for (int r = 0; r < 8; ++r) {
      da[r] = hi | a_lbo | (a0 + (r % v.a_tiles) * a_step);
      db[r] = hi | b_lbo | (b0 + (r % v.b_tiles) * b_step);
      dd[r] = tm + (r % v.n_acc) * acc_cols;
    }
    // warm-up
    if (elect_one()) {
      for (int r = 0; r < 64; ++r) umma_f16(tm, hi | a_lbo | a0, hi | b_lbo | b0, idesc, 1);
      tc_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    const long long t0 = clock64();
    if (elect_one()) {
      if (v.table == 3) {
        // conv2's real stage: per channel pair 4 wide (N = 128) + 2 narrow (N = 64) entries, 2 tiles each; optional
        // tcgen05.fence / commits between "stages" of 24 MMAs (v.commit bit 0: commits, bit 1: fence)
        const uint32_t idesc_n = umma_idesc_bf16(128, 64), idesc_w = umma_idesc_bf16(128, 128);
        const uint32_t lb = ((5 * 64u * 16) >> 4) << 16;
        for (int r = 0; r < v.reps; r += 24) {
          const uint32_t unit_lo = a0 + (r & 8), stage_lo = b0 + (r & 16);
          if (v.commit & 2) tc_fence_after();
#pragma unroll
          for (int pr = 0; pr < 2; ++pr)
#pragma unroll
            for (int e = 0; e < 6; ++e) {
              const int q = e < 4 ? e + 1 : (e == 4 ? 0 : 5);
              const bool wide = q >= 1 && q <= 4;
              const uint64_t bdesc = hi | lb | (stage_lo + pr * 640 + (4 - (q == 5 ? 4 : q)) * 64);
              const uint32_t aa = a_lbo | (unit_lo + pr * 512 + (q & 1) * 256 + (q >> 1) * 52);
#pragma unroll
              for (int i = 0; i < 2; ++i)
                if (i < v.n_acc) umma_f16(tm + i * 128 + (q == 5 ? 64 : 0), hi | (aa + i * 128), bdesc, wide ? idesc_w : idesc_n, 1);
            }
          if (v.commit & 1) { tc_commit(&bar2); tc_commit(&bar2); }
        }
      } else if (v.table == 2) {
        // schedule as compile-time offsets from run-time uniform bases (what a per-layer template would generate)
        const uint32_t nb = (v.N * 32) >> 4, acc_stride = v.N / 2;
        const int nt = v.n_acc;
        for (int r = 0; r < v.reps; r += 24) {
          const uint32_t unit_lo = a0 + (r & 8), stage_lo = b0 + (r & 16);   // change per "stage"
#pragma unroll
          for (int j = 0; j < 12; ++j) {
            const uint64_t bdesc = hi | b_lbo | (stage_lo + (j % 4) * nb);
            const uint32_t aa = a_lbo | (unit_lo + (j % 8) * 8 + v.commit);  // commit field reused: A misalignment in 16-byte units
#pragma unroll
            for (int i = 0; i < 2; ++i)
              if (i < nt) umma_f16(tm + i * 2 * acc_stride, hi | (aa + i * 128), bdesc, idesc, 1);
          }
        }
      } else if (v.table) {
        const uint32_t unit_lo = a0, stage_lo = b0, d_base = tm, acc_stride = v.N / 2;
        const int nt = v.n_acc;  // runtime, as in the kernel
        for (int r = 0; r < v.reps; r += 24) {
#pragma unroll 4
          for (int j = 0; j < 12; ++j) {
            const uint4 k4 = s_tab[j];
            const uint64_t bdesc = hi | (k4.y + stage_lo);
            const uint32_t aa = k4.x + unit_lo, d0 = d_base + k4.z;
            const uint32_t acc = (k4.w & 32) ? 0u : 1u;
#pragma unroll
            for (int i = 0; i < 2; ++i)
              if (i < nt) umma_f16(d0 + i * 2 * acc_stride, hi | (aa + i * 128), bdesc, idesc, acc);
          }
        }
      } else
      for (int r = 0; r < v.reps; r += 24) {
#pragma unroll
        for (int u = 0; u < 24; ++u) umma_f16(dd[u & 7], da[u & 7], db[u & 7], idesc, 1);
        if (v.commit) tc_commit(&bar2);
        if (v.delay) {
          const long long t = clock64();
          while (clock64() - t < v.delay) {}
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 1);
    const long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (tid == 0) mbar_arrive(&bar3);
  } else if (tid >= 128 && v.poll) {
    if (v.poll == 1 || (tid & 31) == 0) mbar_wait(&bar3, 0);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long* d_cyc;
  cudaMalloc(&d_cyc, sms * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<Variant> vs;
  for (int layout = 0; layout < 2; ++layout)
    for (int N : {32, 64, 96, 128, 192, 256})
      for (int n_acc : {1, 2, 4}) {
        if (n_acc * N > 512) continue;
        if (layout == 0 && n_acc == 2) vs.push_back(Variant{N, layout, n_acc, 4008, 8, 8, 0, 0, 0, 0});
      }
  for (int N : {64, 128})
    for (int commit : {0, 1})
      for (int delay : {0, 100, 200, 400, 800, 1600}) vs.push_back(Variant{N, 0, 2, 4008, 8, 8, 0, delay, commit, 0});
  for (int N : {64, 128}) vs.push_back(Variant{N, 0, 2, 4008, 8, 8, 1, 0, 0, 0});
  for (int N : {64, 128}) for (int mis : {0, 1, 4, 7}) vs.push_back(Variant{N, 0, 2, 4008, 8, 8, 2, 0, mis, 0});
  for (int c : {0, 1, 2, 3}) vs.push_back(Variant{128, 0, 2, 4008, 8, 8, 3, 0, c, 0});
  // long runs: does the sustained (power-limited) rate differ from the burst rate?
  for (int reps : {40080, 400800, 2004000}) vs.push_back(Variant{128, 0, 2, reps, 8, 8, 3, 0, 0, 0});
  for (int N : {128, 256}) vs.push_back(Variant{N, 0, 2, 2004000, 8, 8, 2, 0, 0, 0});
  for (int poll : {1, 2}) vs.push_back(Variant{128, 0, 2, 40080, 8, 8, 3, 0, 0, poll});
  vs.push_back(Variant{128, 0, 1, 40080, 8, 8, 3, 0, 0, 0});  // one tile per entry: the second MMA of every entry is predicated off (cycles are per SLOT)
  printf("%-8s %4s %5s %6s %6s %14s %12s %12s %10s\n", "layout", "N", "accs", "delay", "commit", "cycles/MMA", "tensor min", "fetch B/cyc", "ms wall");
  for (const Variant& v : vs) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<<<sms, 384, 200 * 1024>>>(v, d_cyc);  // warm
    cudaEventRecord(e0);
    rate_kernel<<<sms, 384, 200 * 1024>>>(v, d_cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(err)); return 1; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (long long c : h) avg += static_cast<double>(c);
    avg /= sms * static_cast<double>(v.reps);
    const double bytes = 128 * 32 + v.N * 32;
    printf("%-8s %4d %5d %6d %6d/%d %12.1f %12.1f %12.1f %10.3f\n", v.table == 3 ? "conv2" : v.table == 2 ? "templ" : v.table ? "table" : (v.layout ? "sw128" : "noswz"), v.N, v.n_acc, v.delay, v.commit, v.poll, avg, v.N / 2.0, bytes / avg, ms);
  }
  return 0;
}
