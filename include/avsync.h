/*
 * avsync.h — C-ABI of the B200-native AV sync-scoring hot path (libavsync_b200.so).
 *
 * The reference (Hu-xiao-max/Alignment-Between-Speech-and-Visual-Mouth-Movements)
 * has no FFI of its own: its seam is plain Python callables.  Each
 * entry point below replaces the arithmetic behind one of those callables; the
 * Python shims in alignment-between-speech-and-visual-mouth-movements_b200/ keep the
 * reference names and signatures and are the only callers.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the ABI.
 *   - every function returns 0 (AVS_OK) or a negative AVS_E* code and never throws;
 *     avs_last_error_string() describes the last failure on the calling thread.
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and the
 *     call returns without synchronising (except the *_host entry points, which
 *     synchronise before returning because their outputs are host buffers).
 *   - device pointers are owned by the caller and must stay alive until the stream
 *     has drained; outputs are fully overwritten.
 *   - sm_100a only: avs_device_check() fails on anything else; there is no CPU path.
 */
#ifndef AVSYNC_H_
#define AVSYNC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVS_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define AVS_API __attribute__((visibility("default")))
#else
#define AVS_API
#endif

enum {
  AVS_OK = 0,
  AVS_EINVAL = -1,     /* bad argument (null pointer, unsupported shape) */
  AVS_ECUDA = -2,      /* a CUDA runtime call failed; see avs_last_error_string() */
  AVS_EARCH = -3,      /* device is not sm_100 */
  AVS_EWORKSPACE = -4, /* workspace too small */
  AVS_ENOMEM = -5
};

/* arithmetic the STCNN / GRU input GEMMs run in */
enum {
  AVS_PREC_FP32 = 0,   /* CUDA-core FFMA, fp32 throughout (slow; bring-up / cross-check path) */
  AVS_PREC_BF16 = 1,   /* tcgen05 kind::f16 bf16 x bf16 -> fp32 accum, single pass */
  AVS_PREC_BF16X3 = 2  /* tcgen05, hi/lo bf16 split, 3 passes (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo): fp32-grade */
};

/* geometry fixed by the reference model (model.py:22-52, misalignment_detection_train.py:79-88) */
#define AVS_T 75
#define AVS_H 50
#define AVS_W 100
#define AVS_EMB 6912      /* 96 * 6 * 12 */
#define AVS_VSTATS 13824  /* 2 * AVS_EMB */
#define AVS_NFFT 2048
#define AVS_NMELS 128

AVS_API int avs_version(void);
/* sha256 (hex) of the sources this binary was compiled from (csrc/*.cu, *.cuh, Makefile, include/avsync.h, in sorted
 * file-name order): lets a loader prove that a prebuilt .so matches the tree it sits in. */
AVS_API const char* avs_source_hash(void);
AVS_API const char* avs_last_error_string(void);
/* 0 iff `device` is compute capability 10.0 (the library is compiled for sm_100a only). */
AVS_API int avs_device_check(int device);

/* ---------------------------------------------------------------- K1: MFCC statistics sweep
 * Replaces compute_audio_stats(shift_audio(audio, k, fps, sr), sr, n_mfcc)
 * (misalignment_detection_train.py:100-127, librosa.feature.mfcc at :121) for all K shifts
 * of B clips in one call.  shift_samples[k] is the signed integer delay shift_audio applies
 * (:103).  The plan de-duplicates STFT frames shared between shifts on the host. */
typedef struct avs_mfcc_plan avs_mfcc_plan;
AVS_API int avs_mfcc_plan_create(int n_samples, int sample_rate, int n_mfcc,
                         const int32_t* shift_samples, int n_shifts, avs_mfcc_plan** out);
AVS_API void avs_mfcc_plan_destroy(avs_mfcc_plan* plan);
/* Host-only (no device needed): the frame plan itself.  frames_out (may be NULL) receives up to
 * n_shifts*n_frames triples (start, lo, hi) — call once with NULL to size it by *n_unique_out —
 * and map_out (may be NULL) the [n_shifts][n_frames] -> unique-frame-id table. */
AVS_API int avs_mfcc_plan_describe(int n_samples, int sample_rate, const int32_t* shift_samples, int n_shifts,
                           int* n_frames_out, int* n_unique_out, int32_t* frames_out, int32_t* map_out);
AVS_API int avs_mfcc_plan_unique_frames(const avs_mfcc_plan* plan); /* distinct STFT frames per clip */
AVS_API int avs_mfcc_plan_frames(const avs_mfcc_plan* plan);        /* STFT frames per shifted signal */
AVS_API size_t avs_mfcc_workspace_bytes(const avs_mfcc_plan* plan, int n_clips);
/* audio: device f32 [n_clips, n_samples]; out_stats: device f32 [n_clips, n_shifts, 2*n_mfcc]
 * = [mean(n_mfcc), unbiased std(n_mfcc)] per (clip, shift). */
AVS_API int avs_mfcc_stats_sweep(const avs_mfcc_plan* plan, const float* audio, int n_clips,
                         float* out_stats, void* workspace, size_t workspace_bytes, void* stream);
/* debug/parity: also write the per-frame MFCC table f32 [n_clips, n_shifts, n_frames, n_mfcc] */
AVS_API int avs_mfcc_sweep_debug(const avs_mfcc_plan* plan, const float* audio, int n_clips,
                         float* out_stats, float* out_mfcc, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---------------------------------------------------------------- K2: STCNN
 * Replaces extract_visual_embeddings (misalignment_detection_train.py:130-144) == the
 * conv half of LipNet.forward (model.py:67-82), eval mode, plus the visual statistics of
 * misalignment_detection_train.py:165.  Weights are the reference's state_dict tensors
 * (device f32, OIDHW): conv1 [32,1,3,5,5], conv2 [64,32,3,5,5], conv3 [96,64,3,3,3]. */
typedef struct avs_stcnn avs_stcnn;
AVS_API int avs_stcnn_create(const float* w1, const float* b1, const float* w2, const float* b2,
                     const float* w3, const float* b3, int precision, void* stream,
                     avs_stcnn** out);
AVS_API void avs_stcnn_destroy(avs_stcnn* net);
AVS_API size_t avs_stcnn_workspace_bytes(const avs_stcnn* net, int n_clips);
/* frames: device f32 [n_clips,1,75,50,100].  out_emb (f32 [n_clips,75,6912], feature index
 * c*72+h*12+w) and out_vstats (f32 [n_clips,13824] = [mean_t, unbiased std_t]) may each be NULL. */
AVS_API int avs_stcnn_forward(const avs_stcnn* net, const float* frames, int n_clips, float* out_emb,
                      float* out_vstats, void* workspace, size_t workspace_bytes, void* stream);
/* Same, from the 8-bit pixels the reference's frames tensor is made of (dataset.py:226-231: frames = float32(u8 / 255.0)):
 * frames device u8 [n_clips,1,75,50,100].  Bit-identical outputs to avs_stcnn_forward on the f32 tensor of those pixels. */
AVS_API int avs_stcnn_forward_u8(const avs_stcnn* net, const uint8_t* frames, int n_clips, float* out_emb,
                         float* out_vstats, void* workspace, size_t workspace_bytes, void* stream);
/* debug/parity: copy the pooled activations of layer 1/2 out as f32 NCDHW
 * ([n,32,75,25,50] / [n,64,75,12,25]); either pointer may be NULL. */
AVS_API int avs_stcnn_forward_debug(const avs_stcnn* net, const float* frames, int n_clips, float* out_emb,
                            float* out_vstats, float* out_pool1, float* out_pool2, void* workspace,
                            size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- K3: Bi-GRU head
 * Replaces gru1 -> gru2 -> fc -> log_softmax of LipNet.forward (model.py:84-95), eval mode.
 * Weight pointers are the reference state_dict tensors, device f32:
 *   w_ih [2][3H, in], w_hh [2][3H, H], b_ih [2][3H], b_hh [2][3H]  ([0]=forward, [1]=reverse),
 *   gate order (r, z, n), H = hidden.  fc_w [V, 2H], fc_b [V]. */
typedef struct avs_bigru avs_bigru;
AVS_API int avs_bigru_create(int in_dim, int hidden, int vocab,
                     const float* g1_w_ih, const float* g1_w_hh, const float* g1_b_ih, const float* g1_b_hh,
                     const float* g2_w_ih, const float* g2_w_hh, const float* g2_b_ih, const float* g2_b_hh,
                     const float* fc_w, const float* fc_b, int precision, void* stream, avs_bigru** out);
AVS_API void avs_bigru_destroy(avs_bigru* head);
AVS_API size_t avs_bigru_workspace_bytes(const avs_bigru* head, int n_clips, int n_steps);
/* emb: device f32 [n_clips, n_steps, in_dim]; out_logp: device f32 [n_clips, n_steps, vocab]. */
AVS_API int avs_bigru_forward(const avs_bigru* head, const float* emb, int n_clips, int n_steps,
                      float* out_logp, void* workspace, size_t workspace_bytes, void* stream);

/* fp32-grade GEMM on tcgen05 (the K3 input-projection kernel, exposed for unit tests and for the
 * DFT-as-GEMM comparison in profiles/): C[M,N] = A[M,K] . W[N,K]^T + bias[N], all device f32 row-major,
 * operands split hi/lo in bf16 (three products, fp32 accumulate).  K % 32 == 0, N % 4 == 0. */
AVS_API size_t avs_gemm_split_workspace_bytes(int M, int N, int K);
AVS_API int avs_gemm_split(const float* a, const float* w, const float* bias, float* c, int M, int N, int K,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- K4: shift-sweep score
 * Replaces, for every (clip, shift), torch.sigmoid(MisalignmentDetector(cat[vstats, astats_k]))
 * (misalignment_detection_train.py:207,243-250,267; misalignment_detection_demo.py:249-250) and
 * the arg-max over shifts.  w1 [hidden, v_dim + a_dim] row-major (classifier.0.weight), b1 [hidden],
 * w2 [hidden] (classifier.3.weight), b2 [1]; all device f32.
 * out_scores f32 [n_clips, n_shifts]; out_best int32 [n_clips] (first maximum). */
AVS_API size_t avs_sweep_score_workspace_bytes(int n_clips, int hidden);
AVS_API int avs_sweep_score(const float* vstats, const float* astats, int n_clips, int n_shifts, int v_dim,
                    int a_dim, const float* w1, const float* b1, const float* w2, const float* b2,
                    int hidden, float* out_scores, int32_t* out_best, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- K5: greedy CTC decode
 * Replaces the arithmetic of decode_prediction (utils.py:20-30): per step arg-max over the vocab
 * (first index on ties), drop repeats and blanks.  logp device f32 [n_clips, n_steps, vocab];
 * out_ids int32 [n_clips, n_steps] (first out_len[i] entries valid, rest -1); out_len int32 [n_clips]. */
AVS_API int avs_ctc_greedy(const float* logp, int n_clips, int n_steps, int vocab, int blank,
                   int32_t* out_ids, int32_t* out_len, void* stream);

/* ---------------------------------------------------------------- whole sweep (K2 + K1 + K4)
 * The composed +-S sync sweep of SURVEY.md §3.2: for each clip, visual stats once, audio stats
 * per shift, detector score per shift, arg-max.  The handle owns its device workspace, two
 * streams and (for the host entry point) pinned staging buffers; clips are processed in chunks
 * of at most `chunk_clips`. */
typedef struct avs_sweep avs_sweep;
AVS_API int avs_sweep_create(const avs_stcnn* net, const avs_mfcc_plan* plan, const float* w1,
                     const float* b1, const float* w2, const float* b2, int hidden,
                     int chunk_clips, avs_sweep** out);
AVS_API void avs_sweep_destroy(avs_sweep* sw);
/* frames/audio/out_* are DEVICE buffers; enqueues on `stream`, no host or device synchronisation (per-call buffers grow
 * with stream-ordered allocations).  Calls on one handle are serialised on the device in issue order, whatever streams
 * they are issued on (each call waits for the handle's previous one), so run / run_host may be mixed freely; calls on one
 * handle must not be issued concurrently from several host threads. */
AVS_API int avs_sweep_run(avs_sweep* sw, const float* frames, const float* audio, int n_clips,
                  float* out_scores, int32_t* out_best, void* stream);
/* frames/audio/out_* are HOST buffers (pageable or pinned); copies are pipelined against
 * compute chunk by chunk on the handle's own streams; returns after the results are in the host buffers.  The detector
 * and STCNN weights the handle points to must be complete (their producing streams synchronised) before the call. */
AVS_API int avs_sweep_run_host(avs_sweep* sw, const float* frames_host, const float* audio_host,
                       int n_clips, float* out_scores_host, int32_t* out_best_host);
/* u8-pixel variants (frames u8 [n_clips,1,75,50,100], 375 KB per clip instead of 1.5 MB): scores bit-identical to the f32
 * entry points on float32(u8 / 255.0) frames. */
AVS_API int avs_sweep_run_u8(avs_sweep* sw, const uint8_t* frames, const float* audio, int n_clips,
                     float* out_scores, int32_t* out_best, void* stream);
AVS_API int avs_sweep_run_host_u8(avs_sweep* sw, const uint8_t* frames_host, const float* audio_host,
                          int n_clips, float* out_scores_host, int32_t* out_best_host);
/* ---------------------------------------------------------------- decode metrics
 * Edit distances behind calculate_cer / calculate_wer (train.py:945-993) and the positional character
 * matches of evaluate_model (utils.py:83-86) for a batch of id sequences that are already on the device
 * (e.g. straight from avs_ctc_greedy).  pad_id renders as the five characters "<pad>" like the reference
 * table (p_id/a_id/d_id = ids of 'p','a','d'); words are runs of non-space symbols.
 * out: device i32 [n_clips, 6] = {char distance, target chars, word distance, target words,
 * positional matches, predicted chars}.  CER = out[0]/out[1], WER = out[2]/out[3]. */
AVS_API int avs_edit_metrics(const int32_t* pred_ids, const int32_t* pred_len, int pred_stride,
                     const int32_t* tgt_ids, const int32_t* tgt_len, int tgt_stride, int n_clips,
                     int space_id, int pad_id, int p_id, int a_id, int d_id, int32_t* out, void* stream);

/* ---------------------------------------------------------------- pre-processing prologue
 * The per-frame arithmetic of GridDataset.process_video (dataset.py:209-254) for decoded uint8 frames:
 * BGR->gray, crop [0.6h:, 0.3w:0.7w], bilinear resize to 100x50, /255, pad/truncate to 75 frames.
 * Bit-exact with OpenCV's 8-bit cvtColor / resize when the crop has >= 50 rows.
 * frames: device u8 [n_clips, n_frames_in, h, w, channels]; lengths (device i32 [n_clips], may be NULL):
 * valid frames per clip; out: device f32 [n_clips, 1, 75, 50, 100]. */
typedef struct avs_preproc avs_preproc;
AVS_API int avs_preproc_create(int h, int w, int channels, avs_preproc** out);
AVS_API void avs_preproc_destroy(avs_preproc* p);
AVS_API int avs_preproc_crop(const avs_preproc* p, int* y0, int* x0, int* crop_h, int* crop_w);
AVS_API int avs_preproc_run(const avs_preproc* p, const uint8_t* frames, int n_clips, int n_frames_in,
                    const int32_t* lengths, float* out, void* stream);

/* ---------------------------------------------------------------- sample-rate conversion (second half)
 * Replaces librosa.resample(audio, orig_sr, target_sr) in FeatureExtractor.build_feature
 * (misalignment_detection_train.py:202-204) for clips whose audio is not at cfg.sample_rate.  Band-limited polyphase
 * interpolation with a Kaiser-windowed sinc (64 zero crossings, roll-off 0.9476, beta 14.77: resampy's "kaiser_best", the
 * filter librosa shipped before soxr), in the formulation of torchaudio.functional.resample; librosa's current default
 * (soxr_hq) is an un-vendored C library and stays unpinned.
 * in: device f32 [n_signals, n_in]; out: device f32 [n_signals, avs_resample_out_len(plan, n_in)] = ceil(n_in * target / orig). */
typedef struct avs_resample_plan avs_resample_plan;
AVS_API int avs_resample_plan_create(int orig_sr, int target_sr, avs_resample_plan** out);
AVS_API void avs_resample_plan_destroy(avs_resample_plan* plan);
AVS_API long long avs_resample_out_len(const avs_resample_plan* plan, long long n_in);
AVS_API int avs_resample(const avs_resample_plan* plan, const float* in, long long n_in, int n_signals, float* out,
                 void* stream);

/* Optional per-kernel timing: when enabled, every launch of a profiled kernel is bracketed by CUDA
 * events on its own stream.  slot: 0 pack, 1 conv1, 2 conv2, 3 conv3, 4 vstats, 5 mfcc log-mel,
 * 6 mfcc stats, 7 score GEMM, 8 score, 9 GRU operand pack, 10 GRU input GEMM, 11 GRU recurrence,
 * 12 fc + log_softmax.  avs_prof_read synchronises on the recorded events. */
AVS_API void avs_prof_enable(int on);
AVS_API void avs_prof_reset(void);
AVS_API int avs_prof_read(int slot, double* total_ms, int* count);
#ifdef AVS_EXPERIMENTS
/* Tools build only (make -C csrc EXPERIMENTS=1 -> libavsync_b200_exp.so; never part of the product library).
 * Experiment switches for the tcgen05 conv kernel (results become garbage; used by tools/conv_microbench.py and
 * tools/conv_issuer_split.py to attribute time): 1 = weight stages loaded once, 2 = activation planes loaded once,
 * 4 = epilogue off, 8 = every MMA issued twice, 16 = clock64 split of the issuer warps (printed by block 0),
 * 32 = epilogue reads TMEM only, 64 = epilogue without stores.  0 = normal operation.  The same build reads the
 * environment knobs AVS_CONV{1,2,3}_WSTAGES / _RING, AVS_AUDIO_MODE, AVS_HOST_FIRST_CHUNK, AVS_GRU_FMA. */
AVS_API void avs_debug_set(int flags);
/* begin / end of the idx-th profiled launch of a slot, in ms since the first profiled pack launch */
AVS_API int avs_prof_read_span(int slot, int idx, double* begin_ms, double* end_ms);
#endif
/* How the persistent conv kernels split their work (host mirror of the device code, no GPU needed): the items of a
 * launch, in (clip, tile set, time step) order, are cut into n_ctas contiguous spans of near-equal cost (cost of an item =
 * its tile count; the last tile set of a plane may be partial).  Returns the half-open item range [first, last) of `cta`.
 * n_tiles = 128-position tiles per plane, tiles_per_item = tiles one work item covers. */
AVS_API int avs_conv_item_span(int n_clips, int n_steps, int n_tiles, int tiles_per_item, int n_ctas, int cta, int* first,
                               int* last);
/* number of kernel launches this library has enqueued since load (all handles, this process) */
AVS_API long long avs_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* AVSYNC_H_ */
