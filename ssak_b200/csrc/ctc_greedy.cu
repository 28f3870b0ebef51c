// ctc_greedy.cu -- greedy CTC decode: per-frame argmax (first maximal index), collapse
// consecutive repeats, drop blank.
//
// Replaces torch.argmax(.., -1) + collapse at ssak/infer/general.py:112 (SpeechBrain
// ctc_greedy_decode), ssak/infer/general.py:118 and ssak/infer/transformers_infer.py:84-85
// (argmax + the tokenizer's group-by collapse).  HBM-bound: every emission is read exactly
// once with coalesced (128-bit when aligned) loads; a group of G lanes owns one frame row.
#include "common.cuh"

namespace ssak {

template <int G>
__global__ void __launch_bounds__(256) greedy_argmax_kernel(const float *__restrict__ probs,
                                                            int64_t rows, int64_t T, int V,
                                                            int64_t sb, int64_t st,
                                                            int32_t *__restrict__ ids) {
    const int sub = threadIdx.x % G;
    const int64_t row0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t row_stride = (int64_t)gridDim.x * blockDim.x / G;
    // warp-uniform trip count (the row of the warp's FIRST group decides): the full-mask shuffles below must be
    // executed by all 32 lanes even when the last groups of the warp have run out of rows; those skip the loads
    const int64_t grp = (threadIdx.x & 31) / G;
    for (int64_t row = row0; row - grp < rows; row += row_stride) {
        const bool live = row < rows;
        const int64_t b = live ? row / T : 0, t = live ? row - b * T : 0;
        const float *x = probs + b * sb + t * st;
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (V % 4 == 0);
        if (!live) {
        } else if (vec) {
            const float4 *x4 = reinterpret_cast<const float4 *>(x);
            for (int i = sub; i < V / 4; i += G) {
                const float4 q = __ldg(x4 + i);
                const int c0 = 4 * i;
                if (q.x > bv) { bv = q.x; bi = c0; }
                if (q.y > bv) { bv = q.y; bi = c0 + 1; }
                if (q.z > bv) { bv = q.z; bi = c0 + 2; }
                if (q.w > bv) { bv = q.w; bi = c0 + 3; }
            }
        } else {
            for (int i = sub; i < V; i += G) {
                const float q = __ldg(x + i);
                if (q > bv) { bv = q; bi = i; }
            }
        }
        // group reduction: larger value wins, equal values -> smaller index (first maximum)
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, d, G);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, d, G);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (sub == 0 && live) ids[row] = bi == 0x7fffffff ? 0 : bi;  // all -inf / NaN row -> index 0
    }
}

// One CTA per utterance: keep[t] = id[t] != id[t-1] && id[t] != blank && t < n; compaction by a
// block-wide scan over tiles of 1024 frames.
__global__ void __launch_bounds__(1024) greedy_collapse_kernel(const int32_t *__restrict__ ids,
                                                               int64_t T,
                                                               const int32_t *__restrict__ n_frames,
                                                               int blank,
                                                               int32_t *__restrict__ out,
                                                               int32_t *__restrict__ out_len) {
    __shared__ int warp_sums[32];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t *id = ids + (int64_t)b * T;
    int32_t *o = out + (int64_t)b * T;
    int n = n_frames ? n_frames[b] : (int)T;
    n = n < 0 ? 0 : (n > (int)T ? (int)T : n);
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int t0 = 0; t0 < n; t0 += 1024) {
        const int t = t0 + tid;
        int cur = -1, keep = 0;
        if (t < n) {
            cur = id[t];
            const int prev = t > 0 ? id[t - 1] : -1;
            keep = (cur != prev) && (cur != blank);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int wrank = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_sums[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            int v = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += y;
            }
            warp_sums[lane] = v;  // inclusive
        }
        __syncthreads();
        const int base = s_base + (warp > 0 ? warp_sums[warp - 1] : 0);
        if (keep) o[base + wrank] = cur;
        __syncthreads();
        if (tid == 0) s_base += warp_sums[31];
        __syncthreads();
    }
    for (int t = s_base + tid; t < (int)T; t += 1024) o[t] = -1;  // padding
    if (tid == 0) out_len[b] = s_base;
}

}  // namespace ssak

using namespace ssak;

extern "C" int ssak_ctc_greedy(const float *probs, int64_t B, int64_t T, int64_t V,
                               int64_t stride_b, int64_t stride_t, const int32_t *n_frames,
                               int32_t blank, int32_t *frame_ids, int32_t *out_tokens,
                               int32_t *out_lengths, ssak_stream_t stream) {
    if (!probs || !frame_ids || B <= 0 || T < 0 || V <= 0 || V > (1 << 24))
        return SSAK_ERR_INVALID_ARGUMENT;
    if ((out_tokens == nullptr) != (out_lengths == nullptr)) return SSAK_ERR_INVALID_ARGUMENT;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int64_t rows = B * T;
    if (rows > 0) {
        // lanes per row: enough 16-byte loads in flight per row, no idle lanes on short rows
        int G = 32;
        while (G > 1 && V < 4 * G * 2) G >>= 1;
        const int threads = 256;
        int64_t blocks = (rows * G + threads - 1) / threads;
        const int64_t cap = (int64_t)device_sm_count() * 16;
        if (blocks > cap) blocks = cap;
        switch (G) {
#define SSAK_G(GG) case GG: greedy_argmax_kernel<GG><<<(unsigned)blocks, threads, 0, s>>>(probs, rows, T, (int)V, stride_b, stride_t, frame_ids); break;
            SSAK_G(1) SSAK_G(2) SSAK_G(4) SSAK_G(8) SSAK_G(16) SSAK_G(32)
#undef SSAK_G
        }
        int rc = check_launch();
        if (rc != SSAK_OK) return rc;
    }
    if (out_tokens) {
        greedy_collapse_kernel<<<(unsigned)B, 1024, 0, s>>>(frame_ids, T, n_frames, blank,
                                                            out_tokens, out_lengths);
        return check_launch();
    }
    return SSAK_OK;
}
