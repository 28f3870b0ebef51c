// host_api.cu -- host-buffer entry points of the C ABI (ssak_*_host) and the context that
// owns their device scratch and stream.  These are what a caller without torch binds, and the
// end-to-end benchmark path: HOST pointers in and out, host<->device copies inside the call.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

constexpr int kMaxSub = 16;  // utterance sub-batches pipelined on separate streams (copy/compute overlap)

struct ssak_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t sub[kMaxSub] = {};
    cudaEvent_t ready = nullptr;
    char *scratch = nullptr;
    size_t scratch_bytes = 0;
};

namespace ssak {

static int cuda_fail(cudaError_t e) {
    set_last_cuda_error(e);
    return SSAK_ERR_CUDA;
}
#define SSAK_CUDA(x)                                  \
    do {                                              \
        cudaError_t e_ = (x);                         \
        if (e_ != cudaSuccess) return cuda_fail(e_);  \
    } while (0)

// bump allocator over the context's grow-only device scratch
struct Arena {
    size_t off = 0;
    size_t take(size_t bytes) {
        const size_t o = off;
        off += align_up(bytes, 256);
        return o;
    }
};

static int ensure_scratch(ssak_context *ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return SSAK_OK;
    if (ctx->scratch) {
        SSAK_CUDA(cudaStreamSynchronize(ctx->stream));
        SSAK_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    const size_t want = bytes + bytes / 8;
    SSAK_CUDA(cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return SSAK_OK;
}

}  // namespace ssak

using namespace ssak;

extern "C" int ssak_context_create(int device, ssak_context_t **out) {
    if (!out) return SSAK_ERR_INVALID_ARGUMENT;
    SSAK_CUDA(cudaSetDevice(device));
    ssak_context *ctx = new ssak_context();
    ctx->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return cuda_fail(e);
    }
    for (int i = 0; i < kMaxSub; ++i) {
        e = cudaStreamCreateWithFlags(&ctx->sub[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) { ssak_context_destroy(ctx); return cuda_fail(e); }
    }
    e = cudaEventCreateWithFlags(&ctx->ready, cudaEventDisableTiming);
    if (e != cudaSuccess) { ssak_context_destroy(ctx); return cuda_fail(e); }
    *out = ctx;
    return SSAK_OK;
}

extern "C" void ssak_context_destroy(ssak_context_t *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    for (int i = 0; i < kMaxSub; ++i)
        if (ctx->sub[i]) {
            cudaStreamSynchronize(ctx->sub[i]);
            cudaStreamDestroy(ctx->sub[i]);
        }
    if (ctx->ready) cudaEventDestroy(ctx->ready);
    if (ctx->scratch) cudaFree(ctx->scratch);
    delete ctx;
}

extern "C" int ssak_ctc_loss_host(ssak_context_t *ctx, const float *log_probs_host, int64_t T,
                                  int64_t B, int64_t V, const int32_t *targets_host, int64_t Smax,
                                  const int32_t *input_lengths_host,
                                  const int32_t *target_lengths_host, int32_t blank,
                                  int32_t zero_infinity, const float *grad_out_host,
                                  float *nll_host, float *grad_host) {
    if (!ctx || !log_probs_host || !targets_host || !input_lengths_host || !target_lengths_host ||
        !nll_host || T < 0 || B <= 0 || V <= 0 || Smax < 0)
        return SSAK_ERR_INVALID_ARGUMENT;
    SSAK_CUDA(cudaSetDevice(ctx->device));
    for (int64_t b = 0; b < B; ++b)
        if (input_lengths_host[b] < 0 || input_lengths_host[b] > T || target_lengths_host[b] < 0 ||
            target_lengths_host[b] > Smax)
            return SSAK_ERR_INVALID_ARGUMENT;
    if (blank < 0 || blank >= V) return SSAK_ERR_INVALID_ARGUMENT;
    for (int64_t b = 0; b < B; ++b)   // labels outside the vocabulary are an argument error, never clamped
        for (int64_t i = 0; i < target_lengths_host[b]; ++i)
            if (targets_host[b * Smax + i] < 0 || targets_host[b * Smax + i] >= V) return SSAK_ERR_INVALID_ARGUMENT;
    const bool want_grad = grad_host != nullptr;
    // Utterance sub-batches on separate streams: the host->device copy of sub-batch i+1, the kernels of
    // sub-batch i and the device->host copy of sub-batch i-1 overlap (utterances are independent).
    // (measured on B200, PCIe 5: 19 MB [C2] 4 sub-batches 0.95 ms, 8: 0.95; 309 MB [1k] 4: 9.8 ms, 8: 8.0, 16: 8.5;
    //  1.57 GB [C5] 4: 39 ms, 8: 35.5, 16: 34.1 -- the copies run at ~46 GB/s per direction when both are busy)
    const size_t lp_bytes = (size_t)T * B * V * 4;
    int NS = B >= 32 ? (lp_bytes > ((size_t)1 << 30) ? 16 : (lp_bytes > ((size_t)64 << 20) ? 8 : 4)) : (B >= 8 ? 2 : 1);
    if (const char *e = getenv("SSAK_HOST_SUBBATCHES")) {   // (tuning aid)
        const int v = atoi(e);
        if (v >= 1 && v <= kMaxSub && v <= B) NS = v;
    }
    int64_t b0s[kMaxSub + 1], lmaxs[kMaxSub];
    size_t wsb[kMaxSub], ws_total = 0;
    for (int i = 0; i <= NS; ++i) b0s[i] = B * i / NS;
    for (int i = 0; i < NS; ++i) {
        int64_t lm = 0;
        for (int64_t b = b0s[i]; b < b0s[i + 1]; ++b) lm = std::max<int64_t>(lm, target_lengths_host[b]);
        lmaxs[i] = lm;
        wsb[i] = ssak_ctc_loss_workspace_bytes_v(T, b0s[i + 1] - b0s[i], V, lm, want_grad);
        if (wsb[i] == 0) return SSAK_ERR_UNSUPPORTED;
        ws_total += align_up(wsb[i], 256);
    }
    const size_t n_lp = (size_t)T * B * V;
    Arena a;
    const size_t o_lp = a.take(n_lp * 4), o_grad = a.take(want_grad ? n_lp * 4 : 0),
                 o_tg = a.take((size_t)B * std::max<int64_t>(Smax, 1) * 4), o_off = a.take((size_t)B * 8),
                 o_il = a.take((size_t)B * 4), o_tl = a.take((size_t)B * 4), o_nll = a.take((size_t)B * 4),
                 o_go = a.take((size_t)B * 4), o_ws = a.take(ws_total);
    int rc = ensure_scratch(ctx, a.off);
    if (rc != SSAK_OK) return rc;
    char *d = ctx->scratch;
    cudaStream_t s0 = ctx->stream;
    std::vector<int64_t> offs((size_t)B);
    std::vector<float> ones;
    if (want_grad && !grad_out_host) ones.assign((size_t)B, 1.0f);
    // small arrays first, on the context stream
    if (Smax > 0)
        SSAK_CUDA(cudaMemcpyAsync(d + o_tg, targets_host, (size_t)B * Smax * 4, cudaMemcpyHostToDevice, s0));
    SSAK_CUDA(cudaMemcpyAsync(d + o_il, input_lengths_host, (size_t)B * 4, cudaMemcpyHostToDevice, s0));
    SSAK_CUDA(cudaMemcpyAsync(d + o_tl, target_lengths_host, (size_t)B * 4, cudaMemcpyHostToDevice, s0));
    if (want_grad)
        SSAK_CUDA(cudaMemcpyAsync(d + o_go, grad_out_host ? grad_out_host : ones.data(), (size_t)B * 4,
                                  cudaMemcpyHostToDevice, s0));
    // target offsets are relative to each sub-batch's first row of the padded target matrix
    for (int i = 0; i < NS; ++i)
        for (int64_t b = b0s[i]; b < b0s[i + 1]; ++b) offs[(size_t)b] = (b - b0s[i]) * Smax;
    SSAK_CUDA(cudaMemcpyAsync(d + o_off, offs.data(), (size_t)B * 8, cudaMemcpyHostToDevice, s0));
    SSAK_CUDA(cudaEventRecord(ctx->ready, s0));
    const size_t pitch = (size_t)B * V * 4;
    size_t ws_off = o_ws;
    for (int i = 0; i < NS; ++i) {
        cudaStream_t s = ctx->sub[i];
        const int64_t b0 = b0s[i], nb = b0s[i + 1] - b0s[i];
        SSAK_CUDA(cudaStreamWaitEvent(s, ctx->ready, 0));
        // [T, nb, V] slice of the [T, B, V] host tensor: T rows of nb*V floats
        SSAK_CUDA(cudaMemcpy2DAsync(d + o_lp + (size_t)b0 * V * 4, pitch, log_probs_host + b0 * V, pitch,
                                    (size_t)nb * V * 4, (size_t)T, cudaMemcpyHostToDevice, s));
        const float *lp = (const float *)(d + o_lp) + b0 * V;
        rc = ssak_ctc_loss_forward(lp, T, nb, V, B * V, V, (const int32_t *)(d + o_tg) + b0 * std::max<int64_t>(Smax, 1),
                                   (const int64_t *)(d + o_off) + b0, (const int32_t *)(d + o_il) + b0,
                                   (const int32_t *)(d + o_tl) + b0, lmaxs[i], blank, want_grad ? 1 : 0,
                                   (float *)(d + o_nll) + b0, d + ws_off, wsb[i], s);
        if (rc != SSAK_OK) return rc;
        if (want_grad) {
            float *g = (float *)(d + o_grad) + b0 * V;
            rc = ssak_ctc_loss_backward((const float *)(d + o_go) + b0, lp, T, nb, V, B * V, V,
                                        (const int32_t *)(d + o_tg) + b0 * std::max<int64_t>(Smax, 1),
                                        (const int64_t *)(d + o_off) + b0, (const int32_t *)(d + o_il) + b0,
                                        (const int32_t *)(d + o_tl) + b0, lmaxs[i], blank, zero_infinity,
                                        (const float *)(d + o_nll) + b0, g, B * V, V, d + ws_off, wsb[i], s);
            if (rc != SSAK_OK) return rc;
            SSAK_CUDA(cudaMemcpy2DAsync(grad_host + b0 * V, pitch, g, pitch, (size_t)nb * V * 4, (size_t)T,
                                        cudaMemcpyDeviceToHost, s));
        }
        SSAK_CUDA(cudaMemcpyAsync(nll_host + b0, (float *)(d + o_nll) + b0, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
        ws_off += align_up(wsb[i], 256);
    }
    for (int i = 0; i < NS; ++i) SSAK_CUDA(cudaStreamSynchronize(ctx->sub[i]));
    if (zero_infinity)
        for (int64_t b = 0; b < B; ++b)
            if (std::isinf(nll_host[b]) && nll_host[b] > 0.f) nll_host[b] = 0.f;   // +inf only: NaN stays visible
    return SSAK_OK;
}

extern "C" int ssak_forced_align_host(ssak_context_t *ctx, const float *emissions_host, int64_t B,
                                      int64_t Tmax, int64_t V, const int32_t *tokens_host,
                                      int64_t Lmax, const int32_t *emission_lengths_host,
                                      const int32_t *token_lengths_host, int32_t blank,
                                      int32_t first_as_garbage, const float *col0_host,
                                      int32_t *starts_host, int32_t *ends_host,
                                      double *scores_host, int32_t *t_start_host,
                                      int32_t *status_host) {
    if (!ctx || !emissions_host || !tokens_host || !emission_lengths_host || !token_lengths_host ||
        !starts_host || !ends_host || !scores_host || !t_start_host || !status_host || B <= 0 ||
        Tmax < 0 || V <= 0 || Lmax < 0 || (first_as_garbage && !col0_host))
        return SSAK_ERR_INVALID_ARGUMENT;
    for (int64_t b = 0; b < B; ++b) {
        if (emission_lengths_host[b] < 0 || emission_lengths_host[b] > Tmax || token_lengths_host[b] < 0 ||
            token_lengths_host[b] > Lmax)
            return SSAK_ERR_INVALID_ARGUMENT;
        for (int64_t i = 0; i < token_lengths_host[b]; ++i)
            if (tokens_host[b * Lmax + i] < 0 || tokens_host[b * Lmax + i] >= V) return SSAK_ERR_INVALID_ARGUMENT;
    }
    SSAK_CUDA(cudaSetDevice(ctx->device));
    const size_t n_em = (size_t)B * Tmax * V;
    const size_t ws_bytes = ssak_align_workspace_bytes(B, Tmax, Lmax);
    if (ws_bytes == 0) return SSAK_ERR_UNSUPPORTED;
    const size_t nL = (size_t)B * std::max<int64_t>(Lmax, 1);
    Arena a;
    const size_t o_em = a.take(n_em * 4), o_tk = a.take(nL * 4), o_el = a.take((size_t)B * 4),
                 o_tl = a.take((size_t)B * 4), o_c0 = a.take(first_as_garbage ? (size_t)B * Tmax * 4 : 0),
                 o_st = a.take(nL * 4), o_en = a.take(nL * 4), o_sc = a.take(nL * 8),
                 o_ts = a.take((size_t)B * 4), o_status = a.take((size_t)B * 4), o_ws = a.take(ws_bytes);
    int rc = ensure_scratch(ctx, a.off);
    if (rc != SSAK_OK) return rc;
    char *d = ctx->scratch;
    cudaStream_t s = ctx->stream;
    SSAK_CUDA(cudaMemcpyAsync(d + o_em, emissions_host, n_em * 4, cudaMemcpyHostToDevice, s));
    if (Lmax > 0)
        SSAK_CUDA(cudaMemcpyAsync(d + o_tk, tokens_host, (size_t)B * Lmax * 4, cudaMemcpyHostToDevice, s));
    SSAK_CUDA(cudaMemcpyAsync(d + o_el, emission_lengths_host, (size_t)B * 4, cudaMemcpyHostToDevice, s));
    SSAK_CUDA(cudaMemcpyAsync(d + o_tl, token_lengths_host, (size_t)B * 4, cudaMemcpyHostToDevice, s));
    if (first_as_garbage)
        SSAK_CUDA(cudaMemcpyAsync(d + o_c0, col0_host, (size_t)B * Tmax * 4, cudaMemcpyHostToDevice, s));
    rc = ssak_forced_align((const float *)(d + o_em), B, Tmax, V, Tmax * V, V, (const int32_t *)(d + o_tk),
                           Lmax, Lmax, (const int32_t *)(d + o_el), (const int32_t *)(d + o_tl), blank,
                           first_as_garbage, first_as_garbage ? (const float *)(d + o_c0) : nullptr,
                           (int32_t *)(d + o_st), (int32_t *)(d + o_en), (double *)(d + o_sc),
                           (int32_t *)(d + o_ts), (int32_t *)(d + o_status), nullptr, nullptr, nullptr, d + o_ws,
                           ws_bytes, s);
    if (rc != SSAK_OK) return rc;
    if (Lmax > 0) {
        SSAK_CUDA(cudaMemcpyAsync(starts_host, d + o_st, (size_t)B * Lmax * 4, cudaMemcpyDeviceToHost, s));
        SSAK_CUDA(cudaMemcpyAsync(ends_host, d + o_en, (size_t)B * Lmax * 4, cudaMemcpyDeviceToHost, s));
        SSAK_CUDA(cudaMemcpyAsync(scores_host, d + o_sc, (size_t)B * Lmax * 8, cudaMemcpyDeviceToHost, s));
    }
    SSAK_CUDA(cudaMemcpyAsync(t_start_host, d + o_ts, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    SSAK_CUDA(cudaMemcpyAsync(status_host, d + o_status, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    SSAK_CUDA(cudaStreamSynchronize(s));
    return SSAK_OK;
}

extern "C" int ssak_ctc_greedy_host(ssak_context_t *ctx, const float *probs_host, int64_t B,
                                    int64_t T, int64_t V, const int32_t *n_frames_host,
                                    int32_t blank, int32_t *frame_ids_host,
                                    int32_t *out_tokens_host, int32_t *out_lengths_host) {
    if (!ctx || !probs_host || !out_tokens_host || !out_lengths_host || B <= 0 || T < 0 || V <= 0)
        return SSAK_ERR_INVALID_ARGUMENT;
    SSAK_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)B * T * V, nbt = (size_t)B * std::max<int64_t>(T, 1);
    Arena a;
    const size_t o_p = a.take(n * 4), o_n = a.take((size_t)B * 4), o_id = a.take(nbt * 4),
                 o_out = a.take(nbt * 4), o_len = a.take((size_t)B * 4);
    int rc = ensure_scratch(ctx, a.off);
    if (rc != SSAK_OK) return rc;
    char *d = ctx->scratch;
    cudaStream_t s = ctx->stream;
    SSAK_CUDA(cudaMemcpyAsync(d + o_p, probs_host, n * 4, cudaMemcpyHostToDevice, s));
    if (n_frames_host)
        SSAK_CUDA(cudaMemcpyAsync(d + o_n, n_frames_host, (size_t)B * 4, cudaMemcpyHostToDevice, s));
    rc = ssak_ctc_greedy((const float *)(d + o_p), B, T, V, T * V, V,
                         n_frames_host ? (const int32_t *)(d + o_n) : nullptr, blank, (int32_t *)(d + o_id),
                         (int32_t *)(d + o_out), (int32_t *)(d + o_len), s);
    if (rc != SSAK_OK) return rc;
    if (frame_ids_host && T > 0)
        SSAK_CUDA(cudaMemcpyAsync(frame_ids_host, d + o_id, (size_t)B * T * 4, cudaMemcpyDeviceToHost, s));
    if (T > 0)
        SSAK_CUDA(cudaMemcpyAsync(out_tokens_host, d + o_out, (size_t)B * T * 4, cudaMemcpyDeviceToHost, s));
    SSAK_CUDA(cudaMemcpyAsync(out_lengths_host, d + o_len, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    SSAK_CUDA(cudaStreamSynchronize(s));
    return SSAK_OK;
}
