// K3 — Bi-GRU head of LipNet (model.py:84-95), eval mode: gru1 -> gru2 -> fc -> log_softmax.
//
//   (1) input projections for both directions at once: X[B*T, in] . W_ih[2*3H, in]^T + b_ih  (GEMM)
//   (2) recurrence: one CTA per (clip group, direction), thread j owns hidden unit j; W_hh is kept
//       TRANSPOSED ([k][3H]) so the per-step matrix-vector products read it coalesced; h lives in
//       shared memory; PyTorch gate order (r, z, n), h0 = 0, n = tanh(gi_n + r * (W_hn h + b_hn)).
//   (3) fc GEMM + row-wise log_softmax.
#include "common.cuh"
#include "sgemm.cuh"

struct avs_bigru {
  int in_dim, H, V, precision;
  float* w_ih[2] = {nullptr, nullptr};   // [2*3H, in]
  float* b_ih[2] = {nullptr, nullptr};   // [2*3H]
  float* w_hh_t[2] = {nullptr, nullptr}; // [2][H][3H]  (transposed)
  float* b_hh[2] = {nullptr, nullptr};   // [2*3H]
  float* fc_w = nullptr;                 // [V, 2H]
  float* fc_b = nullptr;
};

namespace avs {

constexpr int kGruClips = 4;  // clips per CTA

// xp: [B*T, 2*3H] (gi for both directions), out: [B, T, 2H]
template <int CB>
__global__ void __launch_bounds__(1024)
gru_recurrence_kernel(const float* __restrict__ xp, const float* __restrict__ w_hh_t, const float* __restrict__ b_hh,
                      float* __restrict__ out, int B, int T, int H) {
  extern __shared__ float s_h[];  // [2][CB][H]
  const int j = threadIdx.x, dir = blockIdx.y, b0 = blockIdx.x * CB;
  const float* wt = w_hh_t + static_cast<size_t>(dir) * H * 3 * H;
  const float br = b_hh[dir * 3 * H + j], bz = b_hh[dir * 3 * H + H + j], bn = b_hh[dir * 3 * H + 2 * H + j];
  float h[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    h[c] = 0.f;
    s_h[c * H + j] = 0.f;
  }
  __syncthreads();
  int cur = 0;
  for (int s = 0; s < T; ++s) {
    const int t = dir ? T - 1 - s : s;
    float gr[CB], gz[CB], gn[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) gr[c] = gz[c] = gn[c] = 0.f;
    const float* hs = s_h + cur * CB * H;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      const float wr = __ldg(wt + static_cast<size_t>(k) * 3 * H + j);
      const float wz = __ldg(wt + static_cast<size_t>(k) * 3 * H + H + j);
      const float wn = __ldg(wt + static_cast<size_t>(k) * 3 * H + 2 * H + j);
#pragma unroll
      for (int c = 0; c < CB; ++c) {
        const float hk = hs[c * H + k];
        gr[c] = fmaf(wr, hk, gr[c]);
        gz[c] = fmaf(wz, hk, gz[c]);
        gn[c] = fmaf(wn, hk, gn[c]);
      }
    }
    float* hn = s_h + (cur ^ 1) * CB * H;
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const int b = b0 + c;
      if (b < B) {
        const float* gi = xp + (static_cast<size_t>(b) * T + t) * 6 * H + dir * 3 * H;
        const float r = 1.f / (1.f + expf(-(gi[j] + gr[c] + br)));
        const float z = 1.f / (1.f + expf(-(gi[H + j] + gz[c] + bz)));
        const float n = tanhf(gi[2 * H + j] + r * (gn[c] + bn));
        h[c] = (1.f - z) * n + z * h[c];
        out[(static_cast<size_t>(b) * T + t) * 2 * H + dir * H + j] = h[c];
      }
      hn[c * H + j] = h[c];
    }
    __syncthreads();
    cur ^= 1;
  }
}

// in-place row-wise log_softmax over V (one warp per row)
__global__ void __launch_bounds__(128)
log_softmax_kernel(float* __restrict__ x, int rows, int V) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = x + static_cast<size_t>(row) * V;
  float mx = -INFINITY;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, r[v]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int v = lane; v < V; v += 32) s += expf(r[v] - mx);
  s = warp_sum(s);
  const float lse = logf(s);
  for (int v = lane; v < V; v += 32) r[v] = (r[v] - mx) - lse;
}

__global__ void transpose_whh_kernel(const float* __restrict__ w, float* __restrict__ wt, int H) {
  // w [2][3H][H] -> wt [2][H][3H]
  const size_t n = static_cast<size_t>(2) * 3 * H * H;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t d = i / (static_cast<size_t>(3) * H * H), rem = i % (static_cast<size_t>(3) * H * H);
    const size_t row = rem / H, k = rem % H;
    wt[d * 3 * H * H + k * 3 * H + row] = w[i];
  }
}

struct GruWs { float* xp; float* o1; float* o2; size_t total; };
static GruWs carve(const avs_bigru* g, int B, int T, void* ws) {
  Carver c(ws);
  GruWs r{};
  r.xp = c.take<float>(static_cast<size_t>(B) * T * 6 * g->H);
  r.o1 = c.take<float>(static_cast<size_t>(B) * T * 2 * g->H);
  r.o2 = c.take<float>(static_cast<size_t>(B) * T * 2 * g->H);
  r.total = align_up(c.off, 256);
  return r;
}

static int dup(float** dst, const float* src, size_t n, cudaStream_t st) {
  AVS_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), n * sizeof(float)));
  AVS_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return AVS_OK;
}

}  // namespace avs

using namespace avs;

extern "C" int avs_bigru_create(int in_dim, int hidden, int vocab, const float* g1_w_ih, const float* g1_w_hh,
                                const float* g1_b_ih, const float* g1_b_hh, const float* g2_w_ih, const float* g2_w_hh,
                                const float* g2_b_ih, const float* g2_b_hh, const float* fc_w, const float* fc_b,
                                int precision, void* stream, avs_bigru** out) {
  AVS_REQUIRE(g1_w_ih && g1_w_hh && g1_b_ih && g1_b_hh && g2_w_ih && g2_w_hh && g2_b_ih && g2_b_hh && fc_w && fc_b && out,
              "null argument");
  AVS_REQUIRE(hidden % 32 == 0 && hidden >= 32 && hidden <= 1024, "hidden must be a multiple of 32 in [32, 1024]");
  AVS_REQUIRE(in_dim % 8 == 0 && vocab > 0, "in_dim must be a multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  avs_bigru* g = new avs_bigru();
  g->in_dim = in_dim; g->H = hidden; g->V = vocab; g->precision = precision;
  const size_t H = hidden;
  const float* wih[2] = {g1_w_ih, g2_w_ih};
  const float* whh[2] = {g1_w_hh, g2_w_hh};
  const float* bih[2] = {g1_b_ih, g2_b_ih};
  const float* bhh[2] = {g1_b_hh, g2_b_hh};
  const size_t ins[2] = {static_cast<size_t>(in_dim), 2 * H};
  int rc = 0;
  for (int l = 0; l < 2 && !rc; ++l) {
    if ((rc = dup(&g->w_ih[l], wih[l], 6 * H * ins[l], st))) break;
    if ((rc = dup(&g->b_ih[l], bih[l], 6 * H, st))) break;
    if ((rc = dup(&g->b_hh[l], bhh[l], 6 * H, st))) break;
    if (cudaMalloc(reinterpret_cast<void**>(&g->w_hh_t[l]), 6 * H * H * sizeof(float)) != cudaSuccess) { rc = AVS_ENOMEM; break; }
    transpose_whh_kernel<<<256, 256, 0, st>>>(whh[l], g->w_hh_t[l], hidden);
    ++g_launches;
  }
  if (!rc) rc = dup(&g->fc_w, fc_w, static_cast<size_t>(vocab) * 2 * H, st);
  if (!rc) rc = dup(&g->fc_b, fc_b, vocab, st);
  if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = AVS_ECUDA;
  if (rc) {
    avs_bigru_destroy(g);
    return rc;
  }
  *out = g;
  return AVS_OK;
}

extern "C" void avs_bigru_destroy(avs_bigru* g) {
  if (!g) return;
  for (int l = 0; l < 2; ++l) {
    cudaFree(g->w_ih[l]); cudaFree(g->b_ih[l]); cudaFree(g->w_hh_t[l]); cudaFree(g->b_hh[l]);
  }
  cudaFree(g->fc_w); cudaFree(g->fc_b);
  delete g;
}

extern "C" size_t avs_bigru_workspace_bytes(const avs_bigru* g, int n_clips, int n_steps) {
  if (!g || n_clips <= 0 || n_steps <= 0) return 0;
  return carve(g, n_clips, n_steps, nullptr).total;
}

extern "C" int avs_bigru_forward(const avs_bigru* g, const float* emb, int B, int T, float* out_logp, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  AVS_REQUIRE(g && emb && out_logp && workspace, "null argument");
  if (B <= 0 || T <= 0) return AVS_OK;
  GruWs w = carve(g, B, T, workspace);
  if (workspace_bytes < w.total) {
    set_error("bigru workspace too small: %zu < %zu", workspace_bytes, w.total);
    return AVS_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int H = g->H, rows = B * T;
  const size_t sm = static_cast<size_t>(2) * kGruClips * H * sizeof(float);
  const float* x = emb;
  float* outs[2] = {w.o1, w.o2};
  const int ins[2] = {g->in_dim, 2 * H};
  int rc;
  for (int l = 0; l < 2; ++l) {
    if ((rc = sgemm_nt(x, ins[l], g->w_ih[l], ins[l], g->b_ih[l], w.xp, 6 * H, rows, 6 * H, ins[l], st))) return rc;
    gru_recurrence_kernel<kGruClips><<<dim3(cdiv(B, kGruClips), 2), H, sm, st>>>(w.xp, g->w_hh_t[l], g->b_hh[l], outs[l], B, T, H);
    AVS_LAUNCHED();
    x = outs[l];
  }
  if ((rc = sgemm_nt(w.o2, 2 * H, g->fc_w, 2 * H, g->fc_b, out_logp, g->V, rows, g->V, 2 * H, st))) return rc;
  log_softmax_kernel<<<cdiv(rows, 4), 128, 0, st>>>(out_logp, rows, g->V);
  AVS_LAUNCHED();
  return AVS_OK;
}
