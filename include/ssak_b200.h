/*
 * ssak_b200.h -- C ABI of libssak_b200.so: the B200 (sm_100a) CTC lattice kernels.
 *
 * This is the drop-in boundary for the ONE hot path of linto-ai/ssak that this library
 * replaces: the batched CTC lattice dynamic program over log-softmax emissions.  The
 * reference is pure Python on top of torch, so "what its FFI would bind" are the torch
 * operators / Python functions cited on each entry point (paths relative to the reference
 * repository unless they start with site-packages/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     allocate, never synchronise and keep no mutable global state (re-entrant from several
 *     host threads and streams, e.g. under nn.DataParallel: one thread per GPU; kernels whose
 *     CTAs depend on each other are launched cooperatively or as clusters, so concurrent calls
 *     on one GPU cannot starve each other);
 *   - the caller owns every buffer, including the workspace whose size the matching
 *     *_workspace_bytes() call returns;
 *   - return value: SSAK_OK (0) or a negative ssak_status_t; ssak_b200_strerror() names it;
 *   - strides are in ELEMENTS; the vocabulary axis is always contiguous (stride 1).
 */
#ifndef SSAK_B200_H_
#define SSAK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SSAK_API
#else
#define SSAK_API __attribute__((visibility("default")))
#endif

typedef void *ssak_stream_t; /* cudaStream_t */

typedef enum {
    SSAK_OK = 0,
    SSAK_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, negative size, blank outside [0,V) ...   */
    SSAK_ERR_UNSUPPORTED = -2,      /* shape outside what the kernels cover (e.g. L too long) */
    SSAK_ERR_WORKSPACE = -3,        /* workspace_bytes smaller than *_workspace_bytes()       */
    SSAK_ERR_CUDA = -4              /* a CUDA runtime call failed (launch error)               */
} ssak_status_t;

SSAK_API int ssak_b200_version(void);            /* major*10000 + minor*100 + patch */
SSAK_API const char *ssak_b200_strerror(int status);
/* cudaError_t of the last failing runtime call on this host thread (0 if none). */
SSAK_API int ssak_b200_last_cuda_error(void);

/* =========================================================================================
 * CTC loss.  Replaces torch.nn.functional.ctc_loss
 *   (site-packages/torch/nn/functional.py:3042-3115 -> aten::_ctc_loss / _ctc_loss_backward)
 * as reached from ssak/train/transformers/wav2vec_train.py:313-325 (via
 * site-packages/transformers/models/wav2vec2/modeling_wav2vec2.py:1727-1736),
 * ssak/train/speechbrain/wav2vec_train.py:66 and ssak/train/nemo/yamls/model.yaml:3.
 *
 * The forward call runs the alpha recursion from the first frame and the beta recursion
 * from the last frame CONCURRENTLY, each over half of the utterance, and joins them in the
 * middle to get the negative log-likelihood; with save_for_backward != 0 the half lattices
 * are kept in the workspace.  The backward call continues both recursions over the other
 * half, fused with the gradient, so every frame is visited twice in total (as in a classic
 * alpha pass + beta pass) but the serial depth of each call is T/2.
 *
 *   log_probs        [T,B,V] fp32, element strides (lp_stride_t, lp_stride_b, 1)
 *   targets          int32, labels of utterance b at targets[target_offsets[b] + i], each in [0,V): a label
 *                    outside the vocabulary makes the utterance's likelihood (and gradient) NaN -- the
 *                    asynchronous device entry points cannot return an error for device data; the
 *                    host-buffer entry points return SSAK_ERR_INVALID_ARGUMENT
 *   target_offsets   int64 [B]    (b*Smax for a padded [B,Smax] tensor, cumsum for 1-D)
 *   input_lengths    int32 [B]    (0 <= . <= T)
 *   target_lengths   int32 [B]    (0 <= . <= max_target_len)
 *   max_target_len   host upper bound on target_lengths (sizes the launch and workspace)
 * Limits: max_target_len <= 4095, T <= 300000 (SSAK_ERR_UNSUPPORTED beyond: the re-centring
 * offsets are kept exact as fp32 integers).
 * ======================================================================================= */

/* Workspace size for forward(+backward).  save_for_backward == 0: join rows only. */
SSAK_API size_t ssak_ctc_loss_workspace_bytes(int64_t T, int64_t B, int64_t max_target_len,
                                              int save_for_backward);
/* The same with the vocabulary size known.  For V <= 128, max_target_len <= 415 and batches that fill the GPU
 * (B >= 1.5 x SM count) the throughput kernels run (one warp per utterance and direction, no stored lattice:
 * checkpoints every 4 frames), and the workspace is ~3x smaller than what the V-agnostic query above must reserve.
 * Either size is accepted by forward / backward. */
SSAK_API size_t ssak_ctc_loss_workspace_bytes_v(int64_t T, int64_t B, int64_t V, int64_t max_target_len,
                                                int save_for_backward);
/* Diagnostics: which kernel family computed each utterance in the last forward (+ backward) call on this workspace:
 * flags_out[b] (device, int32) = 0: throughput kernels; bit 0: recomputed by the log-domain kernels since forward();
 * bit 1: since backward() (the throughput kernels' self-check failed); bit 2: nobody (NaN). */
SSAK_API int ssak_ctc_loss_path_flags(const void *workspace, int64_t T, int64_t B, int64_t V, int64_t max_target_len,
                                      int32_t save_for_backward, int32_t *flags_out, ssak_stream_t stream);
/* Kernels of this library launched by one forward + backward pair of calls on such a shape (bookkeeping for
 * benchmarks; logits != 0: the ssak_ctc_logits_* pair). */
SSAK_API int ssak_ctc_loss_launches(int64_t B, int64_t V, int64_t max_target_len, int32_t logits);

/* 1 when the loss kernels cover the shape, else 0 (max_target_len > 4095, T > 300000, or rows of V floats that do
 * not fit the shared-memory emission ring: V > ~2040).  A caller that replaces a generic operator
 * (ssak_b200.install() over torch.nn.functional.ctc_loss) uses it to delegate what is not covered. */
SSAK_API int ssak_ctc_loss_supported(int64_t T, int64_t B, int64_t V, int64_t max_target_len);
/* 1 when forward(save_for_backward != 0) on such a shape returns a provisional likelihood that the matching backward
 * call finalises (the throughput kernels; see ssak_ctc_loss_forward), else 0. */
SSAK_API int ssak_ctc_loss_nll_is_provisional(int64_t B, int64_t V, int64_t max_target_len);
/* grad[t,b,:] *= per_utterance[b] * scalar_a[0] * scalar_b[0] (device pointers, each may be NULL = 1); rows whose
 * factor is exactly 1 are not touched.  For callers that run backward with a unit upstream gradient right after
 * forward (to finalise the likelihood) and apply the real upstream gradient when autograd delivers it. */
SSAK_API int ssak_ctc_grad_scale(float *grad, int64_t T, int64_t B, int64_t V, int64_t g_stride_t, int64_t g_stride_b,
                                 const float *per_utterance, const float *scalar_a, const float *scalar_b,
                                 ssak_stream_t stream);

/* aten::_ctc_loss(log_probs, targets, input_lengths, target_lengths, blank, zero_infinity)
 *   -> neg_log_likelihood[B] (fp32; +inf for an infeasible utterance -- zero_infinity is
 *      applied by the caller to the loss and by ssak_ctc_loss_backward to the gradient).
 * The second aten output (log_alpha) is replaced by the opaque workspace.
 *   save_for_backward == 0: a forward-only call; always the log-domain kernels (unlimited range).
 *   save_for_backward != 0: where ssak_ctc_loss_nll_is_provisional() says so (the throughput kernels: fp32 block
 *      floating point), neg_log_likelihood is PROVISIONAL until the matching backward call has run on the same
 *      buffers: backward verifies every frame (posterior mass = 1), recomputes the utterances that fail with the
 *      log-domain kernels and REWRITES their neg_log_likelihood[b] (path flag bit 1).  Read the likelihood after
 *      backward (ssak_b200.ctc_loss and ssak_ctc_loss_host do). */
SSAK_API int ssak_ctc_loss_forward(const float *log_probs, int64_t T, int64_t B, int64_t V,
                                   int64_t lp_stride_t, int64_t lp_stride_b,
                                   const int32_t *targets, const int64_t *target_offsets,
                                   const int32_t *input_lengths, const int32_t *target_lengths,
                                   int64_t max_target_len, int32_t blank,
                                   int32_t save_for_backward, float *neg_log_likelihood,
                                   void *workspace, size_t workspace_bytes, ssak_stream_t stream);

/* aten::_ctc_loss_backward(grad, log_probs, targets, input_lengths, target_lengths,
 *                          neg_log_likelihood, log_alpha, blank, zero_infinity) -> [T,B,V]
 *   grad_out     fp32 [B]: upstream gradient per utterance (already including the
 *                reduction's scaling, e.g. 1/(B*clamp(L_b,1)) for 'mean')
 *   grad         fp32 [T,B,V] out, element strides (g_stride_t, g_stride_b, 1); every element
 *                is written: (exp(lp) - posterior) * grad_out[b] for t < input_lengths[b],
 *                0 beyond (torch's convention: the gradient w.r.t. the logits that fed
 *                log_softmax)
 *   neg_log_likelihood  the buffer the matching forward call wrote; entries of utterances whose provisional
 *                likelihood failed the self-check are rewritten (see ssak_ctc_loss_forward)
 *   workspace    the one the matching forward call filled with save_for_backward != 0 */
SSAK_API int ssak_ctc_loss_backward(const float *grad_out, const float *log_probs, int64_t T,
                                    int64_t B, int64_t V, int64_t lp_stride_t,
                                    int64_t lp_stride_b, const int32_t *targets,
                                    const int64_t *target_offsets, const int32_t *input_lengths,
                                    const int32_t *target_lengths, int64_t max_target_len,
                                    int32_t blank, int32_t zero_infinity,
                                    const float *neg_log_likelihood, float *grad,
                                    int64_t g_stride_t, int64_t g_stride_b, void *workspace,
                                    size_t workspace_bytes, ssak_stream_t stream);

/* Same pair on RAW LOGITS (the step before the path): replaces
 *   log_probs = log_softmax(logits, dim=-1) followed by ctc_loss(log_probs, ...)
 *   (site-packages/transformers/models/wav2vec2/modeling_wav2vec2.py:1725-1736, ssak/infer/general.py:99-101,
 *    ssak/train/speechbrain/wav2vec_train.py:54).  The log-probabilities are never written to memory: one
 *   kernel computes the row normalisers [T,B] into the workspace, the lattice kernels form
 *   x - logsumexp(x) on the fly, and `grad` is d loss / d logits (the a-7 formula is that gradient).
 *   Arguments, workspace (ssak_ctc_loss_workspace_bytes) and error codes as above. */
SSAK_API int ssak_ctc_logits_forward(const float *logits, int64_t T, int64_t B, int64_t V,
                                     int64_t stride_t, int64_t stride_b, const int32_t *targets,
                                     const int64_t *target_offsets, const int32_t *input_lengths,
                                     const int32_t *target_lengths, int64_t max_target_len, int32_t blank,
                                     int32_t save_for_backward, float *neg_log_likelihood, void *workspace,
                                     size_t workspace_bytes, ssak_stream_t stream);

SSAK_API int ssak_ctc_logits_backward(const float *grad_out, const float *logits, int64_t T, int64_t B,
                                      int64_t V, int64_t stride_t, int64_t stride_b, const int32_t *targets,
                                      const int64_t *target_offsets, const int32_t *input_lengths,
                                      const int32_t *target_lengths, int64_t max_target_len, int32_t blank,
                                      int32_t zero_infinity, const float *neg_log_likelihood, float *grad,
                                      int64_t g_stride_t, int64_t g_stride_b, void *workspace,
                                      size_t workspace_bytes, ssak_stream_t stream);

/* Reduction of aten::ctc_loss (site-packages/torch/nn/functional.py:3042-3115) fused into one launch:
 *   reduction 0 'none' (loss_out[B]), 1 'mean' = mean_b(nll_b / clamp(L_b,1)), 2 'sum',
 *   3 'mean_volume' = sum_b nll_b / sum_b L_b (NeMo, ssak/train/nemo/yamls/model.yaml:3);
 *   zero_infinity replaces +inf by 0.  grad_scale[B] (may be NULL) receives d loss / d nll_b
 *   (0 for the samples zero_infinity removed), i.e. what grad_out of ssak_ctc_loss_backward is
 *   after multiplication by the upstream scalar gradient. */
SSAK_API int ssak_ctc_loss_reduce(const float *neg_log_likelihood, const int32_t *target_lengths,
                                  int64_t B, int32_t reduction, int32_t zero_infinity, float *loss_out,
                                  float *grad_scale, ssak_stream_t stream);

/* Utterance-sharded loss (ssak_b200/shard.py): the three device-side pieces around the ONE all-reduce
 * of two doubles a multi-GPU step needs (the caller does the collective, e.g. ncclAllReduce on `packed`).
 *   ssak_ctc_shard_pack       packed[0] = local numerator (reduction 1 'mean': sum nll_b / clamp(L_b,1);
 *                             2 'sum' and 3 'mean_volume': sum nll_b), packed[1] = local denominator part
 *                             (B_local / 1 / sum L_b); grad_scale[b] = d numerator / d nll_b (0 where
 *                             zero_infinity dropped the utterance)
 *   ssak_ctc_shard_finish     loss_out[0] = packed[0] / den, inv_den_out[0] = 1 / den, den = global_batch
 *                             ('mean'), 1 ('sum'), max(packed[1], 1) ('mean_volume')
 *   ssak_ctc_shard_grad_scale grad_out[b] = grad_scale[b] * grad_loss[0] * inv_den[0]  (input of
 *                             ssak_ctc_loss_backward) */
SSAK_API int ssak_ctc_shard_pack(const float *neg_log_likelihood, const int32_t *target_lengths, int64_t B,
                                 int32_t reduction, int32_t zero_infinity, double *packed, float *grad_scale,
                                 ssak_stream_t stream);
SSAK_API int ssak_ctc_shard_finish(const double *packed, int32_t reduction, int64_t global_batch, float *loss_out,
                                   float *inv_den_out, ssak_stream_t stream);
SSAK_API int ssak_ctc_shard_grad_scale(const float *grad_scale, const float *grad_loss, const float *inv_den,
                                       int64_t B, float *grad_out, ssak_stream_t stream);

/* =========================================================================================
 * Forced alignment.  Replaces get_trellis + backtrack + merge_repeats
 *   (ssak/utils/align_transcriptions.py:27-70, 79-123, 141-157), i.e. the reference's own
 *   (T+1)x(L+1) max-plus trellis with USE_MAX=True, USE_CHAR_REPEATED=True (:24-25),
 *   the `changed > stayed` strict tie rule (:117) and the first-max end frame (:88),
 * batched over B utterances, as called from compute_alignment (:347,:354,:361).
 *
 *   emissions        [B,Tmax,V] fp32 log-probabilities, strides (em_stride_b, em_stride_t, 1)
 *   tokens           int32 [B,Lmax] (row stride tok_stride), token ids in [0,V)
 *   emission_lengths int32 [B] (T_b <= Tmax), token_lengths int32 [B] (L_b <= Lmax)
 *   col0             optional fp32 [B,Tmax]: column 0 of the trellis for rows 1..T_b when
 *                    first_as_garbage != 0 (:37, computed by the caller with torch ops);
 *                    NULL -> cumulative blank column (:39; fp64 running sum rounded to fp32
 *                    per element, which is what torch.cumsum does on the CPU)
 *   starts, ends     int32 [B,Lmax] out: half-open FRAME span of every token (Segment.start/.end)
 *   scores           fp64 [B,Lmax] out: Segment.score (mean of the per-frame probabilities)
 *   t_start          int32 [B] out: argmax_t trellis[t, L_b] (:88) = number of frames used
 *   status           int32 [B] out: 0 aligned, 1 = the reference's
 *                    RuntimeError("Failed to align (not enough tokens for the duration?)"),
 *                    2 = the call's watchdog fired (a seam poll saw no progress for 10 s: never expected,
 *                    the multi-CTA launch is cooperative / clustered), 3 = a token id outside [0,V)
 *   trellis_dump     optional fp32 [B,Tmax+1,Lmax+1] out: the full trellis (tests only)
 *   path_token       optional int32 [B,Tmax] out: Point.token_index of the frame, -1 off the path
 *   path_prob        optional fp32 [B,Tmax] out: Point.score of the frame (:106-112), 0 off the path
 * ======================================================================================= */
SSAK_API size_t ssak_align_workspace_bytes(int64_t B, int64_t Tmax, int64_t Lmax);

SSAK_API int ssak_forced_align(const float *emissions, int64_t B, int64_t Tmax, int64_t V,
                               int64_t em_stride_b, int64_t em_stride_t, const int32_t *tokens,
                               int64_t tok_stride, int64_t Lmax, const int32_t *emission_lengths,
                               const int32_t *token_lengths, int32_t blank,
                               int32_t first_as_garbage, const float *col0, int32_t *starts,
                               int32_t *ends, double *scores, int32_t *t_start, int32_t *status,
                               float *trellis_dump, int32_t *path_token, float *path_prob,
                               void *workspace, size_t workspace_bytes, ssak_stream_t stream);

/* =========================================================================================
 * Greedy CTC decode.  Replaces torch.argmax(logits, -1) + collapse-repeats + drop-blank
 *   (ssak/infer/general.py:112 -> speechbrain.decoders.ctc_greedy_decode;
 *    ssak/infer/general.py:118, ssak/infer/transformers_infer.py:84-85 -> argmax + the
 *    tokenizer's group-by collapse, site-packages/transformers/models/wav2vec2/
 *    tokenization_wav2vec2.py:307-322).
 *
 *   probs        [B,T,V] fp32, strides (stride_b, stride_t, 1)
 *   n_frames     int32 [B] frames to decode per utterance, or NULL for T
 *   frame_ids    int32 [B,T] out: argmax per frame (first maximal index), every frame
 *   out_tokens   int32 [B,T] out (may be NULL): collapsed, blank-free ids, left-packed
 *   out_lengths  int32 [B] out (may be NULL with out_tokens)
 * ======================================================================================= */
SSAK_API int ssak_ctc_greedy(const float *probs, int64_t B, int64_t T, int64_t V,
                             int64_t stride_b, int64_t stride_t, const int32_t *n_frames,
                             int32_t blank, int32_t *frame_ids, int32_t *out_tokens,
                             int32_t *out_lengths, ssak_stream_t stream);

/* =========================================================================================
 * Host-buffer entry points (what a non-torch caller binds; also the end-to-end benchmark
 * path): same semantics, HOST pointers in and out, host<->device copies and the device
 * scratch are managed by an opaque context bound to one GPU.
 * ======================================================================================= */
typedef struct ssak_context ssak_context_t;
SSAK_API int ssak_context_create(int device, ssak_context_t **out);
SSAK_API void ssak_context_destroy(ssak_context_t *ctx);

/* loss [B] and (optionally, grad_host != NULL) gradient [T,B,V] contiguous for contiguous
 * [T,B,V] host log-probs; targets padded [B,Smax] int32; grad_out_host [B] or NULL (= 1).
 * Pageable or pinned host memory both work (pinned overlaps better). */
SSAK_API int ssak_ctc_loss_host(ssak_context_t *ctx, const float *log_probs_host, int64_t T,
                                int64_t B, int64_t V, const int32_t *targets_host, int64_t Smax,
                                const int32_t *input_lengths_host,
                                const int32_t *target_lengths_host, int32_t blank,
                                int32_t zero_infinity, const float *grad_out_host,
                                float *nll_host, float *grad_host);

SSAK_API int ssak_forced_align_host(ssak_context_t *ctx, const float *emissions_host, int64_t B,
                                    int64_t Tmax, int64_t V, const int32_t *tokens_host,
                                    int64_t Lmax, const int32_t *emission_lengths_host,
                                    const int32_t *token_lengths_host, int32_t blank,
                                    int32_t first_as_garbage, const float *col0_host,
                                    int32_t *starts_host, int32_t *ends_host,
                                    double *scores_host, int32_t *t_start_host,
                                    int32_t *status_host);

SSAK_API int ssak_ctc_greedy_host(ssak_context_t *ctx, const float *probs_host, int64_t B,
                                  int64_t T, int64_t V, const int32_t *n_frames_host,
                                  int32_t blank, int32_t *frame_ids_host,
                                  int32_t *out_tokens_host, int32_t *out_lengths_host);

#ifdef __cplusplus
}
#endif
#endif /* SSAK_B200_H_ */
