// Decode metrics on device id sequences: character / word edit distances behind
// calculate_cer / calculate_wer (train.py:945-993) and the positional character accuracy of
// evaluate_model (utils.py:83-86), for a whole batch without bringing the decoded ids to the host.
//
// Text semantics are kept exactly: an id renders as one character except `pad_id`, which the reference
// table renders as the five characters "<pad>" (dataset.py:43-45) — it is expanded to five symbols
// ('<', 'p', 'a', 'd', '>') before the distance is taken; words are maximal runs of non-space symbols
// (str.split()).  One warp per clip; lane 0 runs the (small) dynamic programmes.
#include "common.cuh"

namespace avs {

constexpr int kMaxSym = 512;   // symbols per sequence after expansion
constexpr int kMaxWords = 256;

__device__ int expand_ids(const int32_t* ids, int n, int pad_id, int p_id, int a_id, int d_id, int16_t* out) {
  int m = 0;
  for (int i = 0; i < n && m + 5 <= kMaxSym; ++i) {
    const int c = ids[i];
    if (c == pad_id) {
      out[m++] = 30001; out[m++] = static_cast<int16_t>(p_id); out[m++] = static_cast<int16_t>(a_id);
      out[m++] = static_cast<int16_t>(d_id); out[m++] = 30002;
    } else {
      out[m++] = static_cast<int16_t>(c);
    }
  }
  return m;
}

__device__ int split_words(const int16_t* s, int n, int space_id, int16_t* start, int16_t* len) {
  int w = 0, i = 0;
  while (i < n && w < kMaxWords) {
    while (i < n && s[i] == space_id) ++i;
    if (i >= n) break;
    const int b = i;
    while (i < n && s[i] != space_id) ++i;
    start[w] = static_cast<int16_t>(b);
    len[w] = static_cast<int16_t>(i - b);
    ++w;
  }
  return w;
}

// out [B][6] = char distance, target chars, word distance, target words, positional matches, predicted chars
__global__ void __launch_bounds__(32)
edit_metrics_kernel(const int32_t* __restrict__ pred, const int32_t* __restrict__ pred_len, int Tp,
                    const int32_t* __restrict__ tgt, const int32_t* __restrict__ tgt_len, int Tt, int space_id,
                    int pad_id, int p_id, int a_id, int d_id, int32_t* __restrict__ out) {
  __shared__ int16_t sp[kMaxSym], st[kMaxSym];
  __shared__ int16_t wps[kMaxWords], wpl[kMaxWords], wts[kMaxWords], wtl[kMaxWords];
  __shared__ int row[kMaxSym + 1];
  if (threadIdx.x != 0) return;
  const int b = blockIdx.x;
  const int m = expand_ids(pred + static_cast<size_t>(b) * Tp, min(max(pred_len[b], 0), Tp), pad_id, p_id, a_id, d_id, sp);
  const int n = expand_ids(tgt + static_cast<size_t>(b) * Tt, min(max(tgt_len[b], 0), Tt), pad_id, p_id, a_id, d_id, st);
  // character Levenshtein (train.py:951-966), one rolling row
  for (int j = 0; j <= n; ++j) row[j] = j;
  for (int i = 1; i <= m; ++i) {
    int diag = row[0];
    row[0] = i;
    for (int j = 1; j <= n; ++j) {
      const int up = row[j];
      row[j] = (sp[i - 1] == st[j - 1]) ? diag : min(min(up, row[j - 1]), diag) + 1;
      diag = up;
    }
  }
  const int cdist = row[n];
  int match = 0;
  for (int i = 0; i < min(m, n); ++i) match += (sp[i] == st[i]);   // utils.py:84 zip(true_text, predicted_text)
  // word Levenshtein (train.py:971-993)
  const int wm = split_words(sp, m, space_id, wps, wpl), wn = split_words(st, n, space_id, wts, wtl);
  for (int j = 0; j <= wn; ++j) row[j] = j;
  for (int i = 1; i <= wm; ++i) {
    int diag = row[0];
    row[0] = i;
    for (int j = 1; j <= wn; ++j) {
      bool eq = wpl[i - 1] == wtl[j - 1];
      for (int k = 0; eq && k < wpl[i - 1]; ++k) eq = sp[wps[i - 1] + k] == st[wts[j - 1] + k];
      const int up = row[j];
      row[j] = eq ? diag : min(min(up, row[j - 1]), diag) + 1;
      diag = up;
    }
  }
  int32_t* o = out + static_cast<size_t>(b) * 6;
  o[0] = cdist; o[1] = n; o[2] = row[wn]; o[3] = wn; o[4] = match; o[5] = m;
}

}  // namespace avs

using namespace avs;

extern "C" int avs_edit_metrics(const int32_t* pred_ids, const int32_t* pred_len, int pred_stride, const int32_t* tgt_ids,
                                const int32_t* tgt_len, int tgt_stride, int n_clips, int space_id, int pad_id, int p_id,
                                int a_id, int d_id, int32_t* out, void* stream) {
  AVS_REQUIRE(pred_ids && pred_len && tgt_ids && tgt_len && out, "null argument");
  AVS_REQUIRE(pred_stride > 0 && tgt_stride > 0 && pred_stride * 5 <= kMaxSym && tgt_stride * 5 <= kMaxSym,
              "sequences longer than 102 ids are not supported");
  if (n_clips <= 0) return AVS_OK;
  edit_metrics_kernel<<<n_clips, 32, 0, static_cast<cudaStream_t>(stream)>>>(pred_ids, pred_len, pred_stride, tgt_ids, tgt_len,
                                                                           tgt_stride, space_id, pad_id, p_id, a_id, d_id, out);
  AVS_LAUNCHED();
  return AVS_OK;
}
