/* asrk.h -- C ABI of the B200-native acoustic-model hot path.
 *
 * Drop-in boundary for the three parts of the reference's data-parallel path
 * (786440445/ASR_DFCNN_Transformer).  The reference has no FFI of its own: the
 * path sits behind plain Python callables.  Each entry
 * point below names the reference callable it replaces; INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - every pointer marked "device" is a CUDA device pointer owned by the caller;
 *     the library allocates nothing persistent.  Scratch memory is caller
 *     provided and sized by the matching *_workspace_bytes() query; it must be
 *     256-byte aligned and is overwritten by every call.
 *   - all work is enqueued on the cudaStream_t argument; no call synchronises.
 *   - every function returns an int status: ASRK_OK or a negative ASRK_E_* code.
 *     Nothing throws, nothing exits.  Per-utterance data problems (infeasible
 *     CTC rows) are reported through a device int32 row_status array so the host
 *     can raise / drop the row exactly as the reference's callers do.
 *   - re-entrant; no global mutable state.
 */
#ifndef ASRK_H_
#define ASRK_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* CUDA stream handle, passed as an opaque pointer (cudaStream_t). */
typedef void* asrk_stream_t;

#define ASRK_OK 0
#define ASRK_E_BADARG (-1)    /* null pointer / negative size / unknown enum            */
#define ASRK_E_SHAPE (-2)     /* a size outside what the kernels support                */
#define ASRK_E_ALIGN (-3)     /* pointer or stride not aligned as documented            */
#define ASRK_E_WORKSPACE (-4) /* workspace too small or misaligned                      */
#define ASRK_E_CUDA (-5)      /* a CUDA runtime call / kernel launch failed             */

/* row_status values written by asrk_ctc_loss_grad_run */
#define ASRK_ROW_OK 0
#define ASRK_ROW_INFEASIBLE 1   /* no valid alignment: loss=+inf, grad=softmax (TF semantics) */
#define ASRK_ROW_NOT_ENOUGH_TIME 2 /* input_len < L + #repeats: TF raises InvalidArgument      */
#define ASRK_ROW_BAD_LENGTH 3   /* input_len < 1 or > T, label_len < 0 or > label_stride,
                                   or a label outside [0,V)                             */
#define ASRK_ROW_NOT_SMALL 4    /* the call carried ASRK_CTC_SMALL_ONLY but this row does not
                                   fit the fused kernel: loss = NaN, gradient rows = 0  */

int asrk_version(void);
const char* asrk_error_string(int code);
/* number of CUDA kernels the library has launched since it was loaded (measurement aid) */
unsigned long long asrk_launch_count(void);

/* ------------------------------------------------------------------------
 * Part 1: spectrogram features (+ optional fused noise mix)
 *   replaces  util/wav_util.py:49-79   compute_fbank            (ASRK_SPEC_FBANK)
 *             util/wav_util.py:82-112  compute_fbank_from_asrt  (ASRK_SPEC_ASRT)
 *             util/noise.py:48-52,108  SNR2K + "signal + K*noise" (noise != NULL)
 *
 * A ragged batch of B mono 16 kHz utterances.  Utterance b owns the
 * sample_counts[b] samples starting at sample_offsets[b] of `samples` (and of
 * `noise`).  For the 16-byte vector path every sample_offsets[b] should be a
 * multiple of 8 (int16) / 4 (float32) samples (pad between utterances); other
 * offsets work through a slower scalar path.  The number of frames of utterance b is frame_offsets[b+1] -
 * frame_offsets[b]; it is decided BY THE HOST with the reference's Python float
 * expression  int(N/fs*1000 - 25)//10 + 1  (wav_util.py:61; asrt: no "+1",
 * :96) and must satisfy 160*(n-1)+400 <= N.  Frame i covers samples
 * [160 i, 160 i + 400) (wav_util.py:67-69), is multiplied by the symmetric
 * Hamming window of :51-52, transformed (400-point DFT, fp64 internally), and
 * bins 0..199 of log(|X| * mag + 1) are produced (mag = 1, asrt: 1/N).
 * ASRK_SPEC_FBANK then z-scores every bin over the frames of the utterance
 * (sklearn.preprocessing.scale semantics, wav_util.py:79).
 *
 * Output row r of utterance b, frame i is  out + (out_row_offsets[b] + i)*200
 * floats (out_row_offsets == NULL: frame_offsets is used, i.e. a ragged
 * [total_frames, 200] matrix; pass b*1600 for the loader's zero-padded
 * [B,1600,200,1] layout of lm_and_am/data_loader.py:107,146).
 *
 * Noise mix (sample_dtype float32 only): when noise != NULL every sample is
 * replaced on the fly by  fl32(signal + fl32(K_b * noise))  (noise.py:108) with
 * K_b = gain[b] if gain != NULL, else computed on the device from snr_db[b]
 * with the float32 arithmetic of noise.py:48-52 (numpy pairwise summation).
 * ------------------------------------------------------------------------ */
#define ASRK_SPEC_FBANK 0       /* compute_fbank: log-magnitude + per-utterance z-score */
#define ASRK_SPEC_ASRT 1        /* compute_fbank_from_asrt: |X|/N, no z-score           */
#define ASRK_SPEC_FBANK_RAW 2   /* compute_fbank before the z-score of wav_util.py:79   */

#define ASRK_DTYPE_I16 0
#define ASRK_DTYPE_F32 1

size_t asrk_spectrogram_workspace_bytes(int batch, long long total_frames);

int asrk_spectrogram_run(const void* samples,               /* device, int16 or float32 [total_samples] */
                         int sample_dtype,                  /* ASRK_DTYPE_*                              */
                         const float* noise,                /* device float32 [total_samples] or NULL    */
                         const float* gain,                 /* device float32 [B] or NULL                */
                         const int* snr_db,                 /* device int32 [B] or NULL                  */
                         const long long* sample_offsets,   /* device int64 [B]  first sample            */
                         const long long* sample_counts,    /* device int64 [B]  number of samples       */
                         const long long* frame_offsets,    /* device int64 [B+1]                        */
                         const long long* out_row_offsets,  /* device int64 [B] or NULL                  */
                         int batch,
                         long long total_frames,            /* = frame_offsets[B] (host copy)            */
                         int mode,                          /* ASRK_SPEC_*                               */
                         float* out,                        /* device float32, rows of 200               */
                         void* workspace, size_t workspace_bytes, asrk_stream_t stream);

/* Mix gains only (noise.py:48-52 SNR2K on the device, float32 numpy semantics):
 * gain_out[b] = fl32(sqrt(es/en)) * fl32(10^(-dB/20)).  Utterances longer than
 * 2^22 samples get a NaN gain (pass explicit gains for those). */
int asrk_snr2k_run(const float* signal, const float* noise, const long long* sample_offsets,
                   const long long* sample_counts, const int* snr_db, int batch, float* gain_out,
                   asrk_stream_t stream);

/* ------------------------------------------------------------------------
 * Part 2: CTC loss forward-backward + gradient w.r.t. the logits
 *   replaces  K.ctc_batch_cost via ctc_lambda, lm_and_am/model/cnn_ctc.py:149-152
 *             tf.nn.ctc_loss_v2(..., blank_index=V-1),
 *                                   lm_and_am/model/acoustic_model2.py:79-80
 *   (TensorFlow CTCLossOp semantics: softmax over the logits inside the op,
 *    ctc_merge_repeated=True, preprocess_collapse_repeated=False.)
 *
 * logits element (t,b,v) is logits[t*stride_t + b*stride_b + v] (V contiguous),
 * so both the TF time-major [T,B,V] and the Keras batch-major [B,T,V] layouts
 * are accepted; grad uses its own strides.  labels is int32 [B, label_stride].
 * label_mode ASRK_LABELS_BY_LENGTH takes labels[b][0..label_len[b]) (Keras
 * ctc_label_dense_to_sparse); ASRK_LABELS_DROP_ZEROS keeps the non-zero entries
 * of the whole row (tf.contrib.layers.dense_to_sparse, acoustic_model2.py:71).
 *   loss[b]   = -log p(labels_b | logits_b)
 *   grad[t,b,v] = grad_scale[b] * (softmax(logits[t,b,:])[v] - occupancy[t,b,v])
 *                 for t < input_len[b], and 0 for t >= input_len[b]
 * grad == NULL computes the loss only.  When tokens != NULL the greedy decode of
 * part 3 is produced from the same read of the logits (merge_repeated = 1).
 * ------------------------------------------------------------------------ */
#define ASRK_LABELS_BY_LENGTH 0
#define ASRK_LABELS_DROP_ZEROS 1

size_t asrk_ctc_workspace_bytes(int T, int B, int label_stride);

/* 1 when an utterance of at most max_input_len frames and max_label_len labels is handled by
 * the fused CTA-per-utterance kernel (32 state pairs, 60 KB of per-frame scalars).  A host that
 * knows the batch maxima (the reference's loader builds input_length / label_length on the host,
 * lm_and_am/data_loader.py:132-148) may then OR ASRK_CTC_SMALL_ONLY into `phases` of
 * asrk_ctc_loss_grad_run_phases: the three generic kernels behind the fused one are not
 * launched at all; a row that breaks the promise is reported as ASRK_ROW_NOT_SMALL. */
int asrk_ctc_fits_fused(int max_input_len, int max_label_len);
#define ASRK_CTC_SMALL_ONLY 0x10000
/* OR into `phases` of asrk_ctc_loss_grad_run_phases: `logits` holds PROBABILITIES p (the softmax output the
 * Keras model hands to K.ctc_batch_cost, lm_and_am/model/cnn_ctc.py:149-152); the op's input is formed on
 * load as log(p + 1e-7) (Keras' epsilon), and `grad` receives the gradient w.r.t. p:
 *   grad[t,b,v] = grad_scale[b] * (y[v] - occupancy[v]) / (p[v] + 1e-7),   y = (p + 1e-7) / sum_v (p + 1e-7)
 * -- the three element-wise passes of the Keras glue (log, its backward, the upstream scale) are inside the
 * kernel.  tokens must be NULL (the greedy decode is defined on the op's input: asrk_ctc_greedy_decode_run). */
#define ASRK_CTC_INPUT_PROB 0x20000

int asrk_ctc_loss_grad_run(const float* logits, long long stride_t, long long stride_b,
                           int T, int B, int V,
                           const int* labels, int label_stride, /* device int32 [B,label_stride] */
                           const int* label_len,                /* device int32 [B]              */
                           const int* input_len,                /* device int32 [B]              */
                           int blank, int label_mode,
                           const float* grad_scale,             /* device float32 [B] or NULL    */
                           float* loss,                         /* device float32 [B]            */
                           float* grad, long long gstride_t, long long gstride_b, /* or NULL     */
                           int* row_status,                     /* device int32 [B]              */
                           int* tokens, int token_stride,       /* device int32 [B,token_stride] or NULL */
                           int* token_len,                      /* device int32 [B] (with tokens) */
                           float* neg_sum_logits,               /* device float32 [B] or NULL     */
                           void* workspace, size_t workspace_bytes, asrk_stream_t stream);

/* K.ctc_batch_cost(y_true, y_pred, input_length, label_length) (lm_and_am/model/cnn_ctc.py:149-152;
 * cnn_rnn_ctc.py:81-84) in one call: y_pred = softmax output, element (t,b,v) at t*stride_t + b*stride_b + v
 * (Keras' [B,T,V]: stride_t = V, stride_b = T*V), blank = V-1, labels masked by label_len, the op's input
 * log(y_pred + 1e-7) formed on load, loss[b] = -log p, grad = d(sum_b grad_scale[b] loss[b]) / d y_pred (or NULL).
 * flags: 0 or ASRK_CTC_SMALL_ONLY. */
int asrk_ctc_batch_cost_run(const float* y_pred, long long stride_t, long long stride_b, int T, int B, int V,
                            const int* labels, int label_stride, const int* label_len, const int* input_len,
                            const float* grad_scale, float* loss, float* grad, long long gstride_t,
                            long long gstride_b, int* row_status, void* workspace, size_t workspace_bytes,
                            asrk_stream_t stream, int flags);

/* Host -> device staging of the logits without their padding.  `src` is the caller's logits in pinned
 * (page-locked, device-mapped) HOST memory, `dst` the device tensor the CTC entry points will read;
 * element (t,b,v) sits at t*stride_t + b*stride_b + v on both sides (so [T,B,V] and [B,T,V] both work,
 * and the two sides may differ).  Only rows t < input_len[b] are transferred -- the padding of a
 * [T,B,V] batch (28 % of an AISHELL-shaped one) never crosses PCIe.  V % 4 == 0, 16-byte aligned. */
int asrk_ctc_stage_logits_run(const float* src, long long src_stride_t, long long src_stride_b,
                              float* dst, long long dst_stride_t, long long dst_stride_b,
                              const int* input_len /* device int32 [B] */, int T, int B, int V,
                              asrk_stream_t stream);

/* The way back: rows t < input_len[b] of the device tensor `src` (the gradient) are written into the caller's
 * pinned, device-mapped HOST tensor `dst` by the SMs; the all-zero rows t >= input_len[b] are not transferred.
 * `stale_len` (device int32 [B], may be NULL): the input lengths of the batch that used `dst` before -- rows
 * input_len[b] <= t < stale_len[b] are set back to zero, and stale_len becomes input_len for the next call (start
 * it at 0 over a cleared buffer).  Without it the host buffer keeps whatever it held in the padding rows.  Same
 * layout rules as the staging call. */
int asrk_ctc_unstage_rows_run(const float* src, long long src_stride_t, long long src_stride_b,
                              float* dst, long long dst_stride_t, long long dst_stride_b,
                              const int* input_len /* device int32 [B] */, int* stale_len, int T, int B, int V,
                              asrk_stream_t stream);

/* Batch reduction feeding the path's only collective (tf.reduce_mean(self.loss),
 * acoustic_model2.py:83): out2[0] = sum of loss[b] over the rows with row_status[b] ==
 * ASRK_ROW_OK (all rows when row_status == NULL), out2[1] = their number, float64,
 * summed in a fixed order.  The caller all-reduces out2 across ranks. */
int asrk_ctc_loss_sum_run(const float* loss, const int* row_status, int B,
                          double* out2, /* device float64 [2] */
                          asrk_stream_t stream);
/* the same, ADDED to out2: a caller that reports the mean every K steps (the reference prints every second
 * step, lm_and_am/train.py:71-73) accumulates on the device and all-reduces once per K steps */
int asrk_ctc_loss_sum_acc_run(const float* loss, const int* row_status, int B, double* out2, asrk_stream_t stream);

/* ------------------------------------------------------------------------
 * Part 3: greedy CTC decode
 *   replaces  tf.nn.ctc_greedy_decoder, lm_and_am/model/acoustic_model2.py:69
 *             (consumed at lm_and_am/test.py:48-52) and
 *             util/utils.py:57-66 decode_ctc (K.ctc_decode(greedy=True))
 * Per frame t < input_len[b] the FIRST maximum over v is taken (strict '>'
 * scan); it is emitted when != blank and (not merge_repeated or != previous
 * frame's class); neg_sum_logits[b] = -sum_t max_v logits.  tokens row b holds
 * token_len[b] ids; the rest of the row is left untouched (the host pads with
 * 0, test.py:51, or -1, Keras).
 * ------------------------------------------------------------------------ */
size_t asrk_ctc_decode_workspace_bytes(int T, int B);

int asrk_ctc_greedy_decode_run(const float* logits, long long stride_t, long long stride_b,
                               int T, int B, int V, const int* input_len, int blank,
                               int merge_repeated,
                               int* tokens, int token_stride, int* token_len,
                               float* neg_sum_logits, /* device float32 [B] or NULL */
                               void* workspace, size_t workspace_bytes, asrk_stream_t stream);

/* ------------------------------------------------------------------------
 * Steps adjacent to the path
 *   asrk_lfr_run            replaces util/utils.py:7-31 build_LFR_features (stack m frames, skip n; the
 *                           tail repeats the last frame), on a ragged batch: utterance b owns input rows
 *                           [in_offsets[b], in_offsets[b+1]) of `dim` floats and output rows
 *                           [out_offsets[b], out_offsets[b+1]) of m*dim floats, out rows = ceil(T_b / n).
 *   asrk_edit_distance_run  replaces tf.edit_distance(decoded, labels) (normalize=True by default),
 *                           lm_and_am/model/acoustic_model2.py:72: Levenshtein distance of hyp[b][0..hyp_len[b])
 *                           against truth[b][0..truth_len[b]) (truth_stride <= 64), divided by truth_len when
 *                           normalize != 0 (empty truth: +inf for a non-empty hypothesis, else 0).
 * ------------------------------------------------------------------------ */
/*   asrk_color_noise_run    replaces util/noise.py:17-34 color_noise after the draw of the normal deviates
 *                           (`normals`, float64, ragged like the PCM: utterance b owns counts[b] values at
 *                           offsets[b]; numpy's global generator is not reproducible on the device, the Python
 *                           surface draws them exactly like the reference): length-N FFT, bins 0..floor(N/2)
 *                           scaled by (k+1)^colour[b], Hermitian rebuild, real inverse FFT, minus the mean,
 *                           divided by the MAXIMUM (not the absolute maximum), float32.  Any N (Bluestein). */
size_t asrk_color_noise_workspace_bytes(int batch, long long max_count);
int asrk_color_noise_run(const double* normals, const long long* offsets, const long long* counts,
                         const double* colour /* device float64 [B] */, int batch, long long max_count,
                         float* out /* device float32, ragged like normals */,
                         void* workspace, size_t workspace_bytes, asrk_stream_t stream);

/*   asrk_logfbank_run       replaces util/wav_util.py:22-31 compute_fbank_from_api =
 *                           python_speech_features.logfbank(signal, fs, nfilt) + sklearn scale (the feature
 *                           call of the live loaders, lm_and_am/data_loader.py:129): float64 samples (as
 *                           soundfile returns them), pre-emphasis `preemph`, rectangular frames of frame_len
 *                           samples at hop frame_step zero-padded at the end (frame counts decided by the host:
 *                           1 + ceil((N - frame_len)/frame_step)), 512-point power spectrum / 512, triangular mel
 *                           filters between the FFT bins mel_bins[0..nfilt+1] (device int32, from the host's
 *                           numpy evaluation of get_filterbanks), log with 0 -> eps; normalise != 0 z-scores
 *                           every filter over the frames of the utterance.  nfilt <= 224, frame_len <= 512. */
int asrk_logfbank_run(const double* samples, const long long* sample_offsets, const long long* sample_counts,
                      const long long* frame_offsets, const long long* out_row_offsets, const int* mel_bins,
                      int batch, long long total_frames, int nfilt, int frame_len, int frame_step,
                      double preemph, int normalise, float* out, asrk_stream_t stream);

int asrk_lfr_run(const float* in, const long long* in_offsets, float* out, const long long* out_offsets,
                 int batch, int dim, int m, int n, long long total_out_rows, asrk_stream_t stream);

int asrk_edit_distance_run(const int* hyp, int hyp_stride, const int* hyp_len,
                           const int* truth, int truth_stride, const int* truth_len,
                           int batch, int normalize, float* out /* device float32 [B] */,
                           asrk_stream_t stream);

/* ------------------------------------------------------------------------
 * Measurement entry points: the same calls with a bit mask of the launches to
 * enqueue, so that a harness can bracket every kernel with its own CUDA events
 * on the launching stream.  ASRK_PHASE_ALL is what the plain entry points pass;
 * phases of one call must be issued in order on one stream with the same
 * arguments and workspace.
 * ------------------------------------------------------------------------ */
#define ASRK_PHASE_SPEC_SETUP 1      /* tables, tile map, (SNR2K gains)            */
#define ASRK_PHASE_SPEC_MAIN 2       /* framing + window + FFT + log-magnitude     */
#define ASRK_PHASE_SPEC_NORMALIZE 4  /* per-utterance z-score (statistics + rows)  */
#define ASRK_PHASE_SPEC_STATS 8      /* without NORMALIZE: the statistics only     */
#define ASRK_PHASE_CTC_PREP 1        /* label lists, feasibility, repeat chains    */
#define ASRK_PHASE_CTC_ROWS 2        /* per-frame log-sum-exp / arg-max / gather   */
#define ASRK_PHASE_CTC_LATTICE 4     /* alpha / beta recursion, loss               */
#define ASRK_PHASE_CTC_GRAD 8        /* gradient rows                              */
#define ASRK_PHASE_CTC_COLLAPSE 16   /* greedy collapse (when tokens != NULL)      */
#define ASRK_PHASE_CTC_FUSED 32      /* fused CTA-per-utterance kernel (small lattices) */
#define ASRK_PHASE_ALL 0xffff
/* asrk_spectrogram_run_phases only: bits 16..30 of `phases` cap the number of CTAs
 * of the persistent spectrogram kernel (0 = one per SM), so that a caller running
 * the CTC kernels on a second stream can leave SMs free for them. */
#define ASRK_SPEC_CTA_LIMIT(n) ((int)(n) << 16)

int asrk_spectrogram_run_phases(const void* samples, int sample_dtype, const float* noise,
                                const float* gain, const int* snr_db,
                                const long long* sample_offsets, const long long* sample_counts,
                                const long long* frame_offsets, const long long* out_row_offsets,
                                int batch, long long total_frames, int mode, float* out,
                                void* workspace, size_t workspace_bytes, asrk_stream_t stream,
                                int phases);

/* The step's tail as ONE kernel: the CTC loss/gradient of a batch bounded to small lattices (flags must carry
 * ASRK_CTC_SMALL_ONLY, see asrk_ctc_fits_fused; ASRK_CTC_INPUT_PROB allowed) with the z-score pass of the FEATURE
 * path as co-work -- every CTA that has finished its utterance, and the CTAs launched beyond the batch, normalise
 * chunks of up to 64 rows of `z_features` in place ((x - mean) / std per utterance and column: wav_util.py:79; the
 * chunks travel as TMA bulk copies through the kernel's shared memory) from the statistics the spectrogram call left
 * in its workspace.  Call order on the feature side:
 *   asrk_spectrogram_run_phases(..., ASRK_PHASE_SPEC_SETUP | ASRK_PHASE_SPEC_MAIN | ASRK_PHASE_SPEC_STATS)
 *   asrk_spectrogram_zscore_handles(workspace, ...) -> z_stats, z_ticket
 * then this call, ordered behind it (same stream or an event).  z_frame_offsets / z_batch / z_total_frames are the
 * spectrogram call's frame_offsets / batch / total_frames; z_row_offsets must be NULL (flat row layout only).
 * Returns ASRK_E_SHAPE -- before anything is launched -- when the batch cannot take the fused kernel (vector path
 * not applicable, not bounded): run ASRK_PHASE_SPEC_NORMALIZE and the plain CTC entry instead. */
int asrk_ctc_loss_grad_zscore_run(const float* logits, long long stride_t, long long stride_b,
                                  int T, int B, int V, const int* labels, int label_stride,
                                  const int* label_len, const int* input_len, int blank, int label_mode,
                                  const float* grad_scale, float* loss, float* grad,
                                  long long grad_stride_t, long long grad_stride_b, int* row_status,
                                  int* tokens, int token_stride, int* token_len, float* neg_sum_logits,
                                  void* workspace, size_t workspace_bytes, asrk_stream_t stream, int flags,
                                  float* z_features, const float* z_stats, const long long* z_frame_offsets,
                                  const long long* z_row_offsets, int z_batch, long long z_total_frames,
                                  int* z_ticket);

/* Where the spectrogram call keeps the per-utterance statistics ([batch][3][200] float32: mean hi, mean lo, 1/std)
 * and the co-work tickets (two ints: next chunk, next utterance; asrk_ctc_loss_grad_zscore_run zeroes them on its
 * stream before it launches) inside its workspace (same batch / total_frames as the run call). */
int asrk_spectrogram_zscore_handles(void* workspace, size_t workspace_bytes, int batch, long long total_frames,
                                    float** stats, int** ticket);

int asrk_ctc_loss_grad_run_phases(const float* logits, long long stride_t, long long stride_b,
                                  int T, int B, int V, const int* labels, int label_stride,
                                  const int* label_len, const int* input_len, int blank,
                                  int label_mode, const float* grad_scale, float* loss, float* grad,
                                  long long gstride_t, long long gstride_b, int* row_status,
                                  int* tokens, int token_stride, int* token_len,
                                  float* neg_sum_logits, void* workspace, size_t workspace_bytes,
                                  asrk_stream_t stream, int phases);

#ifdef __cplusplus
}
#endif
#endif /* ASRK_H_ */
