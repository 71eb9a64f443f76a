/*
 * mas_b200.h -- C ABI of the B200-native (sm_100a) log-prior + Monotonic
 * Alignment Search library, libmas_b200.so.
 *
 * This is the drop-in boundary for the ONE native component of
 * CognitiveModeling/Face-GAN-TTS, model/monotonic_align (a Cython module run
 * on the CPU every training step), plus the torch log-prior block that feeds
 * it.  Each entry point cites the reference interface it replaces; paths are
 * relative to the reference repository root.
 *
 * Conventions
 *   - plain C: raw pointers + sizes, no torch / C++ types.
 *   - every `*_dev` pointer is DEVICE memory owned by the caller; `stream` is
 *     a cudaStream_t passed as void* (NULL = default stream).  Calls are
 *     asynchronous and stream-ordered; nothing synchronises the host.
 *   - process-wide state: only the table of tuning knobs (mas_b200_set_option, read at launch time) and
 *     per-device caches of device properties; calls from different host threads / on different streams do not
 *     interact.  Every launch is self-contained: no kernel of this library waits for another launch.
 *   - return value: MAS_B200_OK or a negative MAS_B200_ERR_* code (argument
 *     errors are detected on the host before anything is launched).  Per-item
 *     data errors that only the device can see (t_x > t_y, t_x < 1: undefined
 *     behaviour in the reference, core.pyx:34) are reported through the
 *     optional `status_dev` array and leave that item's outputs all zero.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point returns MAS_B200_ERR_CUDA.
 */
#ifndef MAS_B200_H_
#define MAS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAS_B200_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------ */
#define MAS_B200_OK               0
#define MAS_B200_ERR_ARG         -1  /* null pointer / non-positive size / bad enum      */
#define MAS_B200_ERR_UNSUPPORTED -2  /* shape outside what the kernels cover             */
#define MAS_B200_ERR_WORKSPACE   -3  /* workspace pointer null or too small              */
#define MAS_B200_ERR_CUDA        -4  /* CUDA runtime error (see mas_b200_last_cuda_error) */
#define MAS_B200_ERR_ALIGN       -5  /* pointer not aligned as the entry point requires  */

/* per-item codes written to status_dev[b] */
#define MAS_B200_ITEM_OK          0
#define MAS_B200_ITEM_BAD_LENGTH  1  /* t_x < 1, t_y < 1, t_x > Tx, t_y > Ty or t_x > t_y */

/* dtype of the dense path output */
#define MAS_B200_PATH_NONE 0   /* do not materialise the dense path            */
#define MAS_B200_PATH_F32  1   /* float32 {0,1}  (what maximum_path returns)    */
#define MAS_B200_PATH_I32  2   /* int32 {0,1}    (what maximum_path_c fills)    */

/* log-prior implementation selector */
#define MAS_B200_LP_AUTO   0
#define MAS_B200_LP_FFMA   1   /* fp32 CUDA-core contraction                     */
#define MAS_B200_LP_TCGEN05 2  /* tcgen05/TMEM 3xTF32 contraction (sm_100a)       */
/* OR-ed into `impl` of mas_b200_log_prior_maximum_path: accepted and ignored (ABI compatibility with round 1). */
#define MAS_B200_WS_PREPARED 0x100

/* default of core.pyx:40 */
#define MAS_B200_MAX_NEG_VAL (-1e9f)

int         mas_b200_abi_version(void);
const char *mas_b200_error_string(int status);
/* cudaError_t of the last failing CUDA call made by the calling thread, 0 if none. */
int         mas_b200_last_cuda_error(void);

/* Bytes of device workspace the calls below need for (B, Tx, Ty).  Holds the
 * per-token [start, duration] table the backtrack emits, and, for shapes whose
 * direction bits (1 bit / cell) do not fit shared memory, the bit planes. */
size_t mas_b200_workspace_bytes(int B, int Tx, int Ty);

/*
 * t_x[b] = sum_x mask[b,x,0], t_y[b] = sum_y mask[b,0,y]  (as int32, truncated).
 * Replaces: model/monotonic_align/__init__.py:18-21 (D2H of the whole mask +
 * numpy sums).  mask_dev: [B,Tx,Ty] float32 contiguous prefix mask.
 */
int mas_b200_lengths_from_mask(const float *mask_dev, int B, int Tx, int Ty,
                               int *t_x_dev, int *t_y_dev, void *stream);

/*
 * Monotonic Alignment Search on device buffers.
 * Replaces: maximum_path_c(paths, values, t_xs, t_ys, max_neg_val)
 *           model/monotonic_align/core.pyx:40-45 (+ maximum_path_each :9-35)
 * and the D2H/H2D bounce around it (model/monotonic_align/__init__.py:16-23).
 *
 *   value_dev   [B,Tx,Ty] float32, element (b,x,y) at b*stride_b + x*stride_x + y
 *               (unit stride along y).  NOT modified (the reference clobbers
 *               its private host copy; here the caller still owns the tensor).
 *               Cells outside [0,t_x) x [0,t_y) are never read, so the
 *               reference's `value * mask` (__init__.py:13) is not needed.
 *   t_x_dev, t_y_dev   [B] int32 true lengths, 1 <= t_x <= t_y.
 *   path_dev    [B,Tx,Ty] contiguous, dtype per path_dtype; fully written
 *               (zeros included); may be NULL with MAS_B200_PATH_NONE.
 *   durations_dev   [B,Tx] int32 = row sums of the path (what
 *               model/face_tts.py:176 recovers by re-reading the dense path);
 *               0 beyond t_x.  May be NULL.
 *   frame_token_dev [B,Ty] int32: the token index of each frame, -1 beyond
 *               t_y.  May be NULL.
 *   status_dev  [B] int32 per-item code, may be NULL.
 *   workspace_dev / workspace_bytes   >= mas_b200_workspace_bytes(B,Tx,Ty),
 *               256-byte aligned.
 * Bit-exact with the reference for every defined input (same fp32 max/add
 * order, same tie-breaking, same -1e9 substitutions, same NaN behaviour).
 */
int mas_b200_maximum_path(const float *value_dev, long long stride_b, long long stride_x,
                          const int *t_x_dev, const int *t_y_dev,
                          int B, int Tx, int Ty, float max_neg_val,
                          void *path_dev, int path_dtype,
                          int *durations_dev, int *frame_token_dev, int *status_dev,
                          void *workspace_dev, size_t workspace_bytes, void *stream);

/*
 * Grad-TTS log-prior, unfused (materialises [B,Tx,Ty]).
 * Replaces: model/face_tts.py:165-171
 *   log_prior[b,x,t] = -0.5*sum_f y[b,f,t]^2 + sum_f mu_x[b,f,x]*y[b,f,t]
 *                      - 0.5*sum_f mu_x[b,f,x]^2 - 0.5*F*log(2*pi)
 * combined as (y_square' + dot) + (mu_square' + const) (the primed terms carry the -0.5).
 *   mu_x_dev [B,F,Tx] float32 contiguous; y_dev [B,F,Ty] float32 contiguous;
 *   log_prior_dev [B,Tx,Ty] float32 contiguous.
 * Within 1e-4 relative of the torch fp32 expression.
 */
int mas_b200_log_prior(const float *mu_x_dev, const float *y_dev,
                       int B, int F, int Tx, int Ty,
                       float *log_prior_dev, int impl, void *stream);

/*
 * Fused log-prior + MAS: mu_x, y -> path / durations / frame_token in one call, entirely on the device.
 * Replaces: model/face_tts.py:165-174 (log-prior block + maximum_path call).
 * ONE kernel (lp_mas_fused.cu): one CTA per utterance for n_feats in {64, 80} with Tx <= 256 or n_feats in
 * {64, 80, 96, 128} with Tx <= 128; a 2-CTA cluster per utterance (one 128-row M-tile each) for Tx in 129..256 when
 * both CTAs of every utterance are resident at once (2B <= SMs; option fused_pair), which also covers n_feats 96 / 128
 * there.  Ty % 4 == 0, 16-byte aligned mu_x / y.  tcgen05 3xTF32 contraction with mu_x parked
 * in tensor memory, its accumulator written tile by tile straight into the shared-memory ring the alignment search
 * reads -- the [Tx,Ty] value matrix never exists in global memory --, then backtrack, durations, frame_token and (if
 * requested) the dense path, whose zeros are streamed out while the search runs.  The path is the bit-exact MAS
 * (core.pyx:9-35 semantics) of the log-prior values the kernel computed (within 1e-4 relative of torch fp32).
 * Other shapes run the serial form: mas_b200_log_prior into the workspace, then mas_b200_maximum_path.
 * workspace: >= mas_b200_fused_workspace_bytes(B,F,Tx,Ty), 256-byte aligned.
 */
size_t mas_b200_fused_workspace_bytes(int B, int F, int Tx, int Ty);
/* No-op kept for ABI compatibility (round 1's two-kernel pipeline kept flags in the workspace; the fused call is ONE
 * kernel now and keeps no state there).  Validates its arguments only. */
int mas_b200_fused_workspace_prepare(void *workspace_dev, size_t workspace_bytes, int B, int F, int Tx, int Ty,
                                     void *stream);
int mas_b200_log_prior_maximum_path(const float *mu_x_dev, const float *y_dev,
                                    const int *t_x_dev, const int *t_y_dev,
                                    int B, int F, int Tx, int Ty, float max_neg_val,
                                    void *path_dev, int path_dtype,
                                    int *durations_dev, int *frame_token_dev, int *status_dev,
                                    void *workspace_dev, size_t workspace_bytes,
                                    int impl, void *stream);

/*
 * Multi-GPU loss bookkeeping: one-sided gather of this rank's durations [B,Tx] into every rank's
 * [world*B, Tx] int32 buffer, rows [rank*B, +B).  peer_ptrs_dev: DEVICE array of `world` pointers, entry r = rank
 * r's buffer as mapped into this device's address space (symmetric memory / cudaIpc / cuMem peer mapping; the
 * kernel stores over NVLink).  Replaces nothing in the reference -- its DDP ranks align their own
 * per_gpu_batchsize shard (config.py:144-145) and never exchange alignments -- but is what a caller needs for global
 * duration statistics; it takes the place of an NCCL all-gather whose host-side cost exceeds the alignment step itself.
 * Completion is stream-ordered on the WRITER; readers synchronise as for any one-sided put (barrier / later collective).
 * The same gather can ride in the fused kernel itself: options peer_dur_ptrs / peer_world / peer_rank.
 */
int mas_b200_put_durations(const int *durations_dev, int B, int Tx, void *const *peer_ptrs_dev, int world, int rank,
                           void *stream);

/*
 * Dense path from integer durations (the inverse op, used at inference).
 * Replaces: generate_path(duration, mask)  model/utils.py:27-40
 *   path[b,x,y] = 1 iff cum[b,x-1] <= y < cum[b,x], x < t_x, y < t_y
 * durations_dev [B,Tx] int32; t_x_dev/t_y_dev [B] int32 (the prefix mask).
 */
int mas_b200_generate_path(const int *durations_dev, const int *t_x_dev, const int *t_y_dev,
                           int B, int Tx, int Ty, void *path_dev, int path_dtype, void *stream);
/*
 * Same with the reference's FLOAT durations (w_ceil * length_scale, model/face_tts.py:118-119,126): the
 * cumulative sum is taken in fp32 in index order, and `t < cum` on integer t (model/utils.py:10) makes token x
 * own the frames [ceil(cum[x-1]), ceil(cum[x])).  Optionally also (or only: path_dtype MAS_B200_PATH_NONE,
 * path_dev NULL) emits the index form frame_token [B,Ty] (-1 where no token), from which
 * mas_b200_gather_mu_y builds the mu_y of model/face_tts.py:128-129 without the dense path.
 */
int mas_b200_generate_path_f32(const float *durations_dev, const int *t_x_dev, const int *t_y_dev,
                               int B, int Tx, int Ty, void *path_dev, int path_dtype,
                               int *frame_token_dev, void *stream);

/* ------------------------------------------------------------------------
 * Consumers of the alignment inside FaceTTS.compute_loss (SURVEY.md section 8
 * rows a1, a6-a8 and f1, f2, f4), driven by the INDEX form of the path that
 * mas_b200_maximum_path emits -- durations [B,Tx], frame_token [B,Ty] and
 * start = exclusive prefix sum of durations -- so the dense [B,Tx,Ty] path is
 * never re-read.  All tensors float32 / int32, contiguous, device memory.
 * Reductions are two-stage in a fixed order: results are deterministic.
 * ------------------------------------------------------------------------ */

/*
 * mask[b,t] = (t < lengths[b]) as float32 0/1, [B,T].
 * Replaces: sequence_mask(length, max_length)  model/utils.py:6-11 (+ the .to(x_mask) cast at
 * model/face_tts.py:161).
 */
int mas_b200_sequence_mask(const int *lengths_dev, int B, int T, float *mask_dev, void *stream);

/*
 * Random-window crop of the mel target and of the path, one launch for the batch.
 * Replaces: the per-utterance Python loop model/face_tts.py:204-211.
 *   cut_len[b] = min(y_lengths[b], out_size)                                  (:205)
 *   y_cut[b,f,t'] = y[b,f,offsets[b]+t']  for t' < cut_len[b], else 0          (:208)
 *   frame_token_cut[b,t'] = frame_token[b,offsets[b]+t'] likewise, else -1     (:209, index form)
 *   cut_mask[b,t'] = (t' < cut_len[b])                                         (:211)
 * offsets_dev [B] int32 is drawn by the caller (the reference draws it with Python's `random`, :188).
 * y [B,F,Ty] -> y_cut [B,F,out_size]; frame_token [B,Ty] -> [B,out_size].  frame_token_dev /
 * frame_token_cut_dev / cut_lengths_dev / cut_mask_dev may be NULL.
 */
int mas_b200_crop_frames(const float *y_dev, const int *frame_token_dev, const int *y_lengths_dev,
                         const int *offsets_dev, int B, int F, int Ty, int out_size,
                         float *y_cut_dev, int *frame_token_cut_dev, int *cut_lengths_dev,
                         float *cut_mask_dev, void *stream);

/*
 * mu_y from the index form of the path.
 * Replaces: mu_y = attn^T @ mu_x^T   model/face_tts.py:217-218 -- a K = Tx GEMM over a one-hot attn that is
 * really the gather mu_y[b,f,t] = mu_x[b,f,frame_token[b,t]] (0 where frame_token < 0, i.e. beyond t_y).
 * The backward is the matching segmented sum over each token's contiguous frames:
 *   grad_mu_x[b,f,x] = sum_{t in [start-off, start-off+dur) and 0 <= t < len} grad_mu_y[b,f,t]
 * with off = offsets_dev[b] (NULL: 0) and len = lengths_dev[b] (NULL: Ty) describing the crop window.
 *   mu_x / grad_mu_x [B,F,Tx], mu_y / grad_mu_y [B,F,Ty] (Ty = out_size when cropped).
 */
int mas_b200_gather_mu_y(const float *mu_x_dev, const int *frame_token_dev, int B, int F, int Tx, int Ty,
                         float *mu_y_dev, void *stream);
int mas_b200_gather_mu_y_backward(const float *grad_mu_y_dev, const int *start_dev, const int *durations_dev,
                                  const int *offsets_dev, const int *lengths_dev,
                                  int B, int F, int Tx, int Ty, float *grad_mu_x_dev, void *stream);

/*
 * Prior loss fused with the mu_y gather.
 * Replaces: model/face_tts.py:217-218 + :233-234
 *   loss = sum_{b,f,t<len[b]} 0.5*((y - mu_y)^2 + log(2*pi)) / (sum_b len[b] * F)
 * mu_y_dev (may be NULL) receives the gathered mu_y for the decoder; loss_dev is one float.
 * Backward (w.r.t. mu_x, both through mu_y): grad_mu_x[b,f,x] =
 *   -grad_loss / (sum len * F) * sum_{t in token x's frames, t < len} (y[b,f,t] - mu_x[b,f,x]).
 * workspace: >= mas_b200_prior_loss_workspace_bytes(B,F,Ty), 8-byte aligned.
 */
size_t mas_b200_prior_loss_workspace_bytes(int B, int F, int Ty);
int mas_b200_prior_loss(const float *y_dev, const float *mu_x_dev, const int *frame_token_dev,
                        const int *y_lengths_dev, int B, int F, int Tx, int Ty,
                        float *mu_y_dev, float *loss_dev, void *workspace_dev, size_t workspace_bytes,
                        void *stream);
int mas_b200_prior_loss_backward(const float *y_dev, const float *mu_x_dev, const int *start_dev,
                                 const int *durations_dev, const int *offsets_dev, const int *y_lengths_dev,
                                 const float *grad_loss_dev, int B, int F, int Tx, int Ty,
                                 float *grad_mu_x_dev, void *stream);

/*
 * Duration loss straight from the integer durations of the backtrack.
 * Replaces: logw_ = log(1e-8 + sum_t attn) * x_mask   model/face_tts.py:176  (a dense re-read of the path)
 *           duration_loss(logw, logw_, x_lengths)      model/utils.py:43-45, call face_tts.py:179
 *   loss = sum_{b,x} (logw[b,x] - logw_[b,x])^2 / sum_b x_lengths[b]
 * logw [B,Tx] float32; logw_target_dev (logw_) and grad_logw_dev (= d loss / d logw) [B,Tx] may be NULL.
 */
int mas_b200_duration_loss(const float *logw_dev, const int *durations_dev, const int *x_lengths_dev,
                           int B, int Tx, float *loss_dev, float *logw_target_dev, float *grad_logw_dev,
                           void *stream);

/*
 * Host -> device transfer of one padded batch that moves only its VALID part over PCIe.
 * Replaces: the `.to(self.device)` copies of FaceTTS.relocate_input (model/face_tts.py:85-89, call :146) for
 * the tensors of this path, when they start in host memory (the e2e measurement; in training mu_x is already
 * on the device).  mu_x [B,F,Tx] / y [B,F,Ty] are padded to the batch maximum (text_encoder.py:417,
 * lrs2_dataset.py:256,265); the kernel reads [0,t_x) / [0,t_y) of every row straight from PAGE-LOCKED host
 * memory (cudaHostAlloc / cudaHostRegister / torch pin_memory; anything else -> MAS_B200_ERR_ARG) and writes
 * the zero padding on the device, so the device tensors equal a plain copy of correctly padded inputs.
 * Also copies the lengths.  Asynchronous on `stream`; the host buffers must stay untouched until it completes.
 * mu_x_pinned may be NULL (mu_x_dev is then untouched): SM-issued zero-copy reads and the copy engine are
 * separate requesters on the PCIe link, so a caller can move the small mu_x with cudaMemcpyAsync on another
 * stream while this kernel pulls y.
 */
int mas_b200_upload_batch(const float *mu_x_pinned, const float *y_pinned,
                          const int *t_xs_pinned, const int *t_ys_pinned,
                          int B, int F, int Tx, int Ty,
                          float *mu_x_dev, float *y_dev, int *t_x_dev, int *t_y_dev, void *stream);

/*
 * HOST-buffer drop-in with the exact argument meaning of the Cython
 * maximum_path_c (model/monotonic_align/core.pyx:40): `paths` int32 [B,Tx,Ty]
 * (overwritten; need not be pre-zeroed), `values` float32 [B,Tx,Ty] (NOT
 * clobbered), t_xs/t_ys int32 [B], all in host memory.  Copies to the current
 * device, runs mas_b200_maximum_path, copies the result back and
 * synchronises.  Returns the number of rejected items (>= 0) or a negative
 * MAS_B200_ERR_*.
 */
int mas_b200_maximum_path_host(int *paths, const float *values,
                               const int *t_xs, const int *t_ys,
                               int B, int Tx, int Ty, float max_neg_val);

/*
 * HOST-buffer fused call: mu_x [B,F,Tx], y [B,F,Ty], t_xs, t_ys in host memory
 * -> durations [B,Tx], frame_token [B,Ty] and (if paths != NULL) the dense
 * int32 path in host memory.  Same return convention as above.
 */
int mas_b200_log_prior_maximum_path_host(const float *mu_x, const float *y,
                                         const int *t_xs, const int *t_ys,
                                         int B, int F, int Tx, int Ty, float max_neg_val,
                                         int *paths, int *durations, int *frame_token);

/*
 * Packed (ragged) batch -- what a collate function that does not pad hands over; replaces the padded H2D copies of
 * relocate_input (model/face_tts.py:85-89) on the host -> device path.  One contiguous buffer
 *   [t_x: B int32][t_y: B int32][pad to 16 bytes][mu: for b, for f: t_x[b] floats][y: for b, for f: t_y[b] floats]
 * crosses PCIe with ONE cudaMemcpyAsync (valid data only); mas_b200_unpack_batch expands it on the device into the
 * zero-padded mu_x [B,F,Tx] / y [B,F,Ty] / length tensors every other entry point takes.
 *   mas_b200_packed_batch_bytes   size of the packed buffer for the given (host) lengths
 *   mas_b200_pack_batch_host      host-side packer for callers that hold padded host tensors (plain memcpy per row)
 *   mas_b200_unpack_batch         device kernel, stream-ordered; packed_dev 16-byte aligned
 */
size_t mas_b200_packed_batch_bytes(const int *t_xs, const int *t_ys, int B, int F);
int mas_b200_pack_batch_host(const float *mu_x, const float *y, const int *t_xs, const int *t_ys, int B, int F, int Tx,
                             int Ty, void *packed, size_t packed_bytes);
int mas_b200_unpack_batch(const void *packed_dev, int B, int F, int Tx, int Ty, float *mu_x_dev, float *y_dev,
                          int *t_x_dev, int *t_y_dev, void *stream);

/* Tuning/diagnostic knobs (not part of the reference interface).
 * mas_b200_set_option("mas_rows_per_lane", R) etc.; returns previous value or
 * INT32_MIN for an unknown key.  Process-wide, read at launch time. */
int mas_b200_set_option(const char *key, int value);
int mas_b200_get_option(const char *key);
/* Diagnostics / tests only: a device pointer option (stored as the two int options key_lo / key_hi).
 *   "mas_debug_ptr"   [B][32] int64  phase stamps of the MAS / fused kernels
 *   "lp_debug_ptr"    [ctas][32] int64 stamps of the unfused tcgen05 log-prior kernel
 *   "fused_dump_ptr"  [B,Tx,Ty] float32: the fused kernel also writes the value tiles its search consumed
 * NULL switches the option off.  Returns MAS_B200_OK or MAS_B200_ERR_ARG for an unknown key. */
int mas_b200_set_pointer_option(const char *key, void *device_ptr);
/* Tests only: launches `ctas` CTAs on `stream` that each hold one whole SM (maximum dynamic shared memory) for `cycles`
 * SM clock cycles -- a foreign kernel that keeps SMs away from the library's kernels (tests/test_gpu_logprior.py). */
int mas_b200_debug_occupy_sms(int ctas, long long cycles, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MAS_B200_H_ */
