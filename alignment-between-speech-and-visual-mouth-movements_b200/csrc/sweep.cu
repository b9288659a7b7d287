// The composed +-S sync sweep: per clip  visual stats once (K2), audio stats per
// shift (K1), detector score per shift + arg-max (K4).  The handle owns device workspaces, a side
// stream for the audio branch and, for the host entry point, pinned staging buffers so that the
// H2D copy of chunk i+1 overlaps the compute of chunk i.
#include <algorithm>
#include "common.cuh"

struct avs_sweep {
  const avs_stcnn* net;
  const avs_mfcc_plan* plan;
  const float *w1, *b1, *w2, *b2;
  int hidden, chunk, K, n_mfcc, n_samples;
  // device buffers (per chunk)
  void* ws_stcnn = nullptr; size_t ws_stcnn_bytes = 0;
  void* ws_mfcc = nullptr;  size_t ws_mfcc_bytes = 0;
  void* ws_score = nullptr; size_t ws_score_bytes = 0;
  float* vstats = nullptr;  // [chunk, 13824]
  float* astats = nullptr;  // [chunk, K, 2*n_mfcc]
  cudaStream_t side = nullptr, copy = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // host entry point: double-buffered device inputs/outputs + pinned staging
  float* d_frames[2] = {nullptr, nullptr};
  float* d_audio[2] = {nullptr, nullptr};
  float* d_scores[2] = {nullptr, nullptr};
  int32_t* d_best[2] = {nullptr, nullptr};
  float* h_frames[2] = {nullptr, nullptr};
  float* h_audio[2] = {nullptr, nullptr};
  float* h_scores[2] = {nullptr, nullptr};
  int32_t* h_best[2] = {nullptr, nullptr};
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  cudaStream_t main = nullptr;
  bool host_ready = false;
};

extern "C" int avs_mfcc_plan_nshifts_internal(const avs_mfcc_plan* p, int* K, int* n_mfcc, int* n_samples);

using namespace avs;

static const size_t kFrameElems = static_cast<size_t>(AVS_T) * AVS_H * AVS_W;

extern "C" int avs_sweep_create(const avs_stcnn* net, const avs_mfcc_plan* plan, const float* w1, const float* b1,
                                const float* w2, const float* b2, int hidden, int chunk_clips, avs_sweep** out) {
  AVS_REQUIRE(net && plan && w1 && b1 && w2 && b2 && out, "null argument");
  AVS_REQUIRE(hidden > 0 && chunk_clips > 0, "bad shape");
  avs_sweep* s = new avs_sweep();
  s->net = net; s->plan = plan; s->w1 = w1; s->b1 = b1; s->w2 = w2; s->b2 = b2;
  s->hidden = hidden; s->chunk = chunk_clips;
  avs_mfcc_plan_nshifts_internal(plan, &s->K, &s->n_mfcc, &s->n_samples);
  s->ws_stcnn_bytes = avs_stcnn_workspace_bytes(net, chunk_clips);
  s->ws_mfcc_bytes = avs_mfcc_workspace_bytes(plan, chunk_clips);
  s->ws_score_bytes = avs_sweep_score_workspace_bytes(chunk_clips, hidden);
  int rc = AVS_OK;
  auto ck = [&](cudaError_t e) { if (e != cudaSuccess && rc == AVS_OK) { set_error("sweep_create: %s", cudaGetErrorString(e)); rc = AVS_ECUDA; } };
  ck(cudaMalloc(&s->ws_stcnn, s->ws_stcnn_bytes));
  ck(cudaMalloc(&s->ws_mfcc, s->ws_mfcc_bytes));
  ck(cudaMalloc(&s->ws_score, s->ws_score_bytes));
  ck(cudaMalloc(reinterpret_cast<void**>(&s->vstats), static_cast<size_t>(chunk_clips) * AVS_VSTATS * sizeof(float)));
  ck(cudaMalloc(reinterpret_cast<void**>(&s->astats), static_cast<size_t>(chunk_clips) * s->K * 2 * s->n_mfcc * sizeof(float)));
  ck(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
  ck(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
  ck(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
  if (rc) {
    avs_sweep_destroy(s);
    return rc;
  }
  *out = s;
  return AVS_OK;
}

extern "C" void avs_sweep_destroy(avs_sweep* s) {
  if (!s) return;
  cudaFree(s->ws_stcnn); cudaFree(s->ws_mfcc); cudaFree(s->ws_score); cudaFree(s->vstats); cudaFree(s->astats);
  for (int i = 0; i < 2; ++i) {
    cudaFree(s->d_frames[i]); cudaFree(s->d_audio[i]); cudaFree(s->d_scores[i]); cudaFree(s->d_best[i]);
    cudaFreeHost(s->h_frames[i]); cudaFreeHost(s->h_audio[i]); cudaFreeHost(s->h_scores[i]); cudaFreeHost(s->h_best[i]);
    if (s->ev_in[i]) cudaEventDestroy(s->ev_in[i]);
    if (s->ev_done[i]) cudaEventDestroy(s->ev_done[i]);
    if (s->ev_out[i]) cudaEventDestroy(s->ev_out[i]);
  }
  if (s->side) cudaStreamDestroy(s->side);
  if (s->copy) cudaStreamDestroy(s->copy);
  if (s->main) cudaStreamDestroy(s->main);
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
  delete s;
}

// one chunk (n <= s->chunk clips), everything device resident; audio branch on the side stream
static int run_chunk(avs_sweep* s, const float* frames, const float* audio, int n, float* scores, int32_t* best,
                     cudaStream_t st) {
  int rc;
  AVS_CUDA(cudaEventRecord(s->ev_fork, st));
  AVS_CUDA(cudaStreamWaitEvent(s->side, s->ev_fork, 0));
  if ((rc = avs_mfcc_stats_sweep(s->plan, audio, n, s->astats, s->ws_mfcc, s->ws_mfcc_bytes, s->side))) return rc;
  AVS_CUDA(cudaEventRecord(s->ev_join, s->side));
  if ((rc = avs_stcnn_forward(s->net, frames, n, nullptr, s->vstats, s->ws_stcnn, s->ws_stcnn_bytes, st))) return rc;
  AVS_CUDA(cudaStreamWaitEvent(st, s->ev_join, 0));
  return avs_sweep_score(s->vstats, s->astats, n, s->K, AVS_VSTATS, 2 * s->n_mfcc, s->w1, s->b1, s->w2, s->b2, s->hidden,
                         scores, best, s->ws_score, s->ws_score_bytes, st);
}

extern "C" int avs_sweep_run(avs_sweep* s, const float* frames, const float* audio, int n_clips, float* out_scores,
                             int32_t* out_best, void* stream) {
  AVS_REQUIRE(s && frames && audio && out_scores && out_best, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int c0 = 0; c0 < n_clips; c0 += s->chunk) {
    const int n = std::min(s->chunk, n_clips - c0);
    int rc = run_chunk(s, frames + c0 * kFrameElems, audio + static_cast<size_t>(c0) * s->n_samples, n,
                       out_scores + static_cast<size_t>(c0) * s->K, out_best + c0, st);
    if (rc) return rc;
  }
  return AVS_OK;
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

static int host_init(avs_sweep* s) {
  if (s->host_ready) return AVS_OK;
  const size_t fb = static_cast<size_t>(s->chunk) * kFrameElems * sizeof(float);
  const size_t ab = static_cast<size_t>(s->chunk) * s->n_samples * sizeof(float);
  const size_t sb = static_cast<size_t>(s->chunk) * s->K * sizeof(float);
  const size_t bb = static_cast<size_t>(s->chunk) * sizeof(int32_t);
  for (int i = 0; i < 2; ++i) {
    AVS_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_frames[i]), fb));
    AVS_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_audio[i]), ab));
    AVS_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_scores[i]), sb));
    AVS_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_best[i]), bb));
    AVS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&s->h_frames[i]), fb, cudaHostAllocDefault));
    AVS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&s->h_audio[i]), ab, cudaHostAllocDefault));
    AVS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&s->h_scores[i]), sb, cudaHostAllocDefault));
    AVS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&s->h_best[i]), bb, cudaHostAllocDefault));
    AVS_CUDA(cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming));
    AVS_CUDA(cudaEventCreateWithFlags(&s->ev_done[i], cudaEventDisableTiming));
    AVS_CUDA(cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming));
  }
  AVS_CUDA(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking));
  AVS_CUDA(cudaStreamCreateWithFlags(&s->main, cudaStreamNonBlocking));
  s->host_ready = true;
  return AVS_OK;
}

extern "C" int avs_sweep_run_host(avs_sweep* s, const float* frames_host, const float* audio_host, int n_clips,
                                  float* out_scores_host, int32_t* out_best_host) {
  AVS_REQUIRE(s && frames_host && audio_host && out_scores_host && out_best_host, "null argument");
  int rc = host_init(s);
  if (rc) return rc;
  const int n_chunks = cdiv(n_clips, s->chunk);
  // page-locked caller buffers are copied from directly; pageable ones go through the pinned staging slots
  const bool direct = is_pinned(frames_host) && is_pinned(audio_host);
  // software pipeline over chunks: stage(i) -> H2D(i) on the copy stream | compute(i) on main | D2H(i) on copy
  for (int i = 0; i < n_chunks; ++i) {
    const int sl = i & 1, c0 = i * s->chunk, n = std::min(s->chunk, n_clips - c0);
    if (i >= 2) {  // slot reuse: results of chunk i-2 must have landed in pinned memory, then drain them
      AVS_CUDA(cudaEventSynchronize(s->ev_out[sl]));
      const int p0 = (i - 2) * s->chunk, pn = std::min(s->chunk, n_clips - p0);
      memcpy(out_scores_host + static_cast<size_t>(p0) * s->K, s->h_scores[sl], static_cast<size_t>(pn) * s->K * sizeof(float));
      memcpy(out_best_host + p0, s->h_best[sl], static_cast<size_t>(pn) * sizeof(int32_t));
    }
    const float* fsrc = frames_host + c0 * kFrameElems;
    const float* asrc = audio_host + static_cast<size_t>(c0) * s->n_samples;
    if (i >= 2) AVS_CUDA(cudaStreamWaitEvent(s->copy, s->ev_done[sl], 0));  // d_frames[sl] is free once chunk i-2 computed
    if (!direct) {
      memcpy(s->h_frames[sl], fsrc, n * kFrameElems * sizeof(float));
      memcpy(s->h_audio[sl], asrc, static_cast<size_t>(n) * s->n_samples * sizeof(float));
      fsrc = s->h_frames[sl];
      asrc = s->h_audio[sl];
    }
    AVS_CUDA(cudaMemcpyAsync(s->d_frames[sl], fsrc, n * kFrameElems * sizeof(float), cudaMemcpyHostToDevice, s->copy));
    AVS_CUDA(cudaMemcpyAsync(s->d_audio[sl], asrc, static_cast<size_t>(n) * s->n_samples * sizeof(float), cudaMemcpyHostToDevice, s->copy));
    AVS_CUDA(cudaEventRecord(s->ev_in[sl], s->copy));
    AVS_CUDA(cudaStreamWaitEvent(s->main, s->ev_in[sl], 0));
    if ((rc = run_chunk(s, s->d_frames[sl], s->d_audio[sl], n, s->d_scores[sl], s->d_best[sl], s->main))) return rc;
    AVS_CUDA(cudaEventRecord(s->ev_done[sl], s->main));
    AVS_CUDA(cudaStreamWaitEvent(s->copy, s->ev_done[sl], 0));
    AVS_CUDA(cudaMemcpyAsync(s->h_scores[sl], s->d_scores[sl], static_cast<size_t>(n) * s->K * sizeof(float), cudaMemcpyDeviceToHost, s->copy));
    AVS_CUDA(cudaMemcpyAsync(s->h_best[sl], s->d_best[sl], static_cast<size_t>(n) * sizeof(int32_t), cudaMemcpyDeviceToHost, s->copy));
    AVS_CUDA(cudaEventRecord(s->ev_out[sl], s->copy));
  }
  for (int i = std::max(0, n_chunks - 2); i < n_chunks; ++i) {
    const int sl = i & 1, c0 = i * s->chunk, n = std::min(s->chunk, n_clips - c0);
    AVS_CUDA(cudaEventSynchronize(s->ev_out[sl]));
    memcpy(out_scores_host + static_cast<size_t>(c0) * s->K, s->h_scores[sl], static_cast<size_t>(n) * s->K * sizeof(float));
    memcpy(out_best_host + c0, s->h_best[sl], static_cast<size_t>(n) * sizeof(int32_t));
  }
  return AVS_OK;
}
