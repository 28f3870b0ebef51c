// ctc_align.cu -- batched forced alignment (the reference's (T+1)x(L+1) max-plus trellis).
//
// Replaces get_trellis + backtrack + merge_repeats of ssak/utils/align_transcriptions.py
// (:27-70, :79-123, :141-157) bit-exactly: same fp32 add/max sequence (no multiplications, so
// no FMA contraction can occur; the file is compiled with -fmad=false as insurance), the fp64
// running sum of the blank column rounded to fp32 per element (:39 on the CPU), the +/-inf
// sentinels (:41-42), the strict `changed > stayed` tie rule (:117) and the first-max end
// frame (:88).
//
// Forward kernel: one CTA per utterance, the L+1 states spread cyclically over the lanes
// (state = warp*32K + k*32 + lane), the trellis row lives in registers, the neighbour state
// comes from one lane rotation per k, emission rows are prefetched by cp.async.bulk into a
// shared-memory ring and gathered at the token columns.  Nothing of the trellis goes to HBM:
// per cell two decision bits (changed > stayed, changed < stayed) are ballot-packed, 32 states
// per word, and streamed out (2 x (L+1)/8 bytes per frame).
// Back-trace kernel: one warp walks the bit matrix from (t_start, L), 32 frames per memory
// round trip (each lane fetches the 32-state window of one frame), then the CTA turns the
// path into per-token frame spans and mean per-frame probabilities (the Segment.score).
#include "common.cuh"

namespace ssak {

struct AlignCfg {
    int K, W, NW;  // states per lane, warps, 32-bit words per bit plane per frame (= K*W)
    int chunk, stages, slot_bytes;
};

struct AlignParams {
    const float *em;
    int64_t B, Tmax;
    int V;
    int64_t sb, st;
    const int32_t *tokens;
    int64_t tok_stride;
    int Lmax;
    const int32_t *em_len, *tok_len;
    int blank, garbage;
    const float *col0;
    uint32_t *bp;        // [B][Tmax][2][NW]
    unsigned char *rec;  // [B][Tmax] decision flags of the frames on the path
    int32_t *starts, *ends, *t_start, *status;
    double *scores;
    float *dump;
    int32_t *path_token;  // optional [B][Tmax]
    float *path_prob;     // optional [B][Tmax]
    AlignCfg cfg;
};

static bool choose_align_cfg(int64_t Lmax, int64_t B, int V, AlignCfg *c) {
    const int64_t P = Lmax + 1;
    int wtarget = (B <= 148) ? 8 : ((B <= 4 * 148) ? 4 : 2);
    const char *s = getenv("SSAK_ALIGN_WARPS");
    if (s && *s) wtarget = atoi(s);
    int K = 0;
    s = getenv("SSAK_ALIGN_K");
    if (s && *s) K = atoi(s);
    if (K == 0) {
        K = 1;
        while (K < 16 && (P + 32 * K - 1) / (32 * K) > wtarget) K *= 2;
    }
    if (K != 1 && K != 2 && K != 4 && K != 8 && K != 16) return false;
    int64_t W = (P + 32 * K - 1) / (32 * K);
    if (W > 32 && K < 8) { K = 8; W = (P + 32 * K - 1) / (32 * K); }
    if (W > (K == 16 ? 16 : 32)) return false;  // L <= 8191
    c->K = K;
    c->W = (int)W;
    c->NW = K * (int)W;
    c->slot_bytes = ring_slot_bytes(V);
    int chunk = 8;
    while (chunk > 1 && chunk * c->slot_bytes > 16384) chunk >>= 1;
    c->chunk = chunk;
    c->stages = 4;
    return true;
}

template <int K>
__global__ void __launch_bounds__(K == 16 ? 512 : 1024, 1) align_forward_kernel(const AlignParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const AlignCfg &c = p.cfg;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);
    const int W = c.W;
    const bool compute = warp < W;
    const int prod_warp = W < 32 ? W : 0;  // dedicated producer warp when the CTA has room

    int Tb = p.em_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.Tmax ? (int)p.Tmax : Tb);
    int L = p.tok_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    if (L == 0 || Tb == 0) {  // reference: empty back-track loop -> "Failed to align"
        if (tid == 0) p.t_start[b] = 0;
        if (p.dump) {  // :41-42 with an empty token list / no frames: column 0 is all +inf
            float *d = p.dump + (int64_t)b * (p.Tmax + 1) * (p.Lmax + 1);
            for (int t = tid; t <= Tb; t += blockDim.x) d[(int64_t)t * (p.Lmax + 1)] = INF;
            if (Tb == 0)
                for (int j = 1 + tid; j <= L; j += blockDim.x) d[j] = -INF;
        }
        return;
    }
    const int V = p.V;
    const float *em_b = p.em + (int64_t)b * p.sb;
    const int32_t *tk = p.tokens + (int64_t)b * p.tok_stride;

    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    float *xchg = reinterpret_cast<float *>(smem + 64);  // [2][32]
    RowRing ring;
    ring.slots = smem + 64 + 256;
    ring.full = full;
    ring.chunk = c.chunk;
    ring.stages = c.stages;
    ring.slot_bytes = c.slot_bytes;
    ring.row_bytes = 4 * V;

    if (tid == 0) {
        for (int s = 0; s < c.stages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }

    // states of this thread, their tokens, trellis row 0 (:35, :41, :42)
    const int jbase = warp * 32 * K + lane;
    int tok[K];
    float v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = jbase + k * 32;
        int tkn = p.blank;
        if (j >= 1 && j <= L) {
            tkn = tk[j - 1];
            tkn = tkn < 0 ? 0 : (tkn >= V ? V - 1 : tkn);
        }
        tok[k] = tkn;
        v[k] = j == 0 ? (L >= Tb + 1 ? INF : 0.f) : -INF;
    }
    const int jL_rel = L - warp * 32 * K;  // state L inside this warp?
    const bool ownsL = compute && jL_rel >= 0 && jL_rel < 32 * K && (jL_rel & 31) == lane;
    const int kL = jL_rel >> 5;
    float best = -INF;  // trellis[0, L] with L >= 1
    int best_t = 0;
    if (p.dump && compute) {
        float *d = p.dump + (int64_t)b * (p.Tmax + 1) * (p.Lmax + 1);
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (jbase + k * 32 <= L) d[jbase + k * 32] = v[k];
    }
    if (compute && lane == 31) xchg[warp] = v[K - 1];
    __syncthreads();

    const int C = c.chunk, NST = c.stages, NW = c.NW;
    RingProducer prod;
    prod.src = em_b;
    prod.step_elems = p.st;
    prod.stage = 0;
    prod.remaining = Tb;
    if (warp == prod_warp && lane == 0)
        for (int n = 0; n < NST; ++n) ring_issue_next(ring, prod);
    int free_at = C;  // iteration at which the oldest in-flight stage has no reader left
    RingPos pos;
    pos.init(em_b, p.st);
    double acc = 0.0;  // :39 cumulative blank column, fp64 running sum (thread 0)
    uint32_t *bp_row = p.bp + (int64_t)b * p.Tmax * 2 * NW;
    const float *col0_b = p.garbage && p.col0 ? p.col0 + (int64_t)b * p.Tmax : nullptr;
    float *dump_row = p.dump ? p.dump + ((int64_t)b * (p.Tmax + 1) + 1) * (p.Lmax + 1) : nullptr;

    for (int t = 0; t < Tb; ++t) {
        const int par = t & 1;
        if (warp == prod_warp && t == free_at) {
            if (lane == 0) ring_issue_next(ring, prod);
            free_at += C;
        }
        if (compute) {
            const float *row = pos.row(ring);
            const float eb = row[p.blank];
            float e[K];
#pragma unroll
            for (int k = 0; k < K; ++k) e[k] = row[tok[k]];
            pos.advance(ring);
            float r[K];
#pragma unroll
            for (int k = 0; k < K; ++k) r[k] = __shfl_sync(FULL, v[k], (lane + 31) & 31);
            const float xin = warp > 0 ? xchg[par * 32 + warp - 1] : 0.f;
            unsigned myword = 0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float prev = lane == 0 ? (k == 0 ? xin : r[k > 0 ? k - 1 : 0]) : r[k];
                const float stayb = v[k] + eb;             // :48
                const float stayt = v[k] + e[k];           // :49
                const float chg = prev + e[k];             // :51
                const float stayed = fmaxf(stayb, stayt);  // what backtrack recomputes (:96-99)
                float nv = fmaxf(stayed, chg);
                bool gt = chg > stayed, lt = chg < stayed;
                if (k == 0 && tid == 0) {  // column 0 (:37 / :39) with the +inf sentinel (:42)
                    if (col0_b) {
                        nv = col0_b[t];
                    } else {
                        acc += (double)eb;
                        nv = (float)acc;
                    }
                    if (t + 1 >= Tb + 1 - L) nv = INF;
                    gt = lt = false;
                }
                v[k] = nv;
                const unsigned bg = __ballot_sync(FULL, gt), bl = __ballot_sync(FULL, lt);
                if (lane == k) myword = bg;
                if (lane == K + k) myword = bl;
            }
            if (lane < 2 * K) bp_row[lane < K ? warp * K + lane : NW + warp * K + (lane - K)] = myword;
            bp_row += 2 * NW;
            if (ownsL) {
                float vl = v[0];
#pragma unroll
                for (int k = 1; k < K; ++k)
                    if (k == kL) vl = v[k];
                if (vl > best) {  // first maximum (:88)
                    best = vl;
                    best_t = t + 1;
                }
            }
            if (dump_row) {
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (jbase + k * 32 <= L) dump_row[jbase + k * 32] = v[k];
                dump_row += p.Lmax + 1;
            }
            if (lane == 31) xchg[(par ^ 1) * 32 + warp] = v[K - 1];
        }
        __syncthreads();
    }
    if (ownsL) p.t_start[b] = best_t;
}

__global__ void __launch_bounds__(128) align_backtrace_kernel(const AlignParams p) {
    __shared__ int s_status;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    int Tb = p.em_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.Tmax ? (int)p.Tmax : Tb);
    int L = p.tok_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    const int NW = p.cfg.NW;
    const uint32_t *bp_b = p.bp + (int64_t)b * p.Tmax * 2 * NW;
    unsigned char *rec = p.rec + (int64_t)b * p.Tmax;
    int32_t *st_b = p.starts + (int64_t)b * p.Lmax;
    int32_t *en_b = p.ends + (int64_t)b * p.Lmax;
    double *sc_b = p.scores + (int64_t)b * p.Lmax;
    const int t_start = (L == 0 || Tb == 0) ? 0 : p.t_start[b];

    if (warp == 0) {
        int t = t_start, j = L;
        bool done = false;
        while (t > 0 && !done) {
            // lane i holds the 32-state window [base, base+32) of trellis row t-i
            const int base = max(j - 31, 0);
            const int i0 = base >> 5, sh = base & 31;
            uint32_t g0 = 0, g1 = 0, l0 = 0, l1 = 0;
            const int rr = t - lane;
            if (rr >= 1) {
                const uint32_t *rowp = bp_b + (int64_t)(rr - 1) * 2 * NW;
                g0 = __ldg(rowp + i0);
                l0 = __ldg(rowp + NW + i0);
                if (i0 + 1 < NW) {
                    g1 = __ldg(rowp + i0 + 1);
                    l1 = __ldg(rowp + NW + i0 + 1);
                }
            }
            const uint32_t wg = __funnelshift_r(g0, g1, sh), wl = __funnelshift_r(l0, l1, sh);
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
                if (t - i < 1) break;
                const uint32_t gi = __shfl_sync(FULL, wg, i), li = __shfl_sync(FULL, wl, i);
                const int bit = j - base;
                const uint32_t gt = (gi >> bit) & 1u, lt = (li >> bit) & 1u;
                const int frame = t - i - 1;
                if (lane == 0) rec[frame] = (unsigned char)(gt | (lt << 1));
                if (gt) {  // :117 changed > stayed -> previous token
                    if (lane == 0) st_b[j - 1] = frame;
                    --j;
                    if (j == 0) { done = true; break; }  // :119-120
                }
            }
            t -= 32;
        }
        if (lane == 0) {
            s_status = done ? 0 : 1;  // :121-122 "Failed to align"
            p.status[b] = done ? 0 : 1;
        }
    }
    __syncthreads();
    const bool ok = s_status == 0;
    int32_t *ptok = p.path_token ? p.path_token + (int64_t)b * p.Tmax : nullptr;
    float *pprob = p.path_prob ? p.path_prob + (int64_t)b * p.Tmax : nullptr;
    if (ptok || pprob) {  // frames off the path (before the first token, after t_start, failures)
        const int lo = ok ? st_b[0] : 0, hi = ok ? t_start : 0;
        for (int f = tid; f < (int)p.Tmax; f += blockDim.x)
            if (f < lo || f >= hi) {
                if (ptok) ptok[f] = -1;
                if (pprob) pprob[f] = 0.f;
            }
    }
    const float *em_b = p.em + (int64_t)b * p.sb;
    const int32_t *tk = p.tokens + (int64_t)b * p.tok_stride;
    for (int i = tid; i < p.Lmax; i += blockDim.x) {
        if (!ok || i >= L) {
            st_b[i] = -1;
            en_b[i] = -1;
            sc_b[i] = 0.0;
            continue;
        }
        // merge_repeats (:141-157): span of token i and the mean of its per-frame scores
        const int s = st_b[i];
        const int e = i + 1 < L ? st_b[i + 1] : t_start;
        int tkn = tk[i];
        tkn = tkn < 0 ? 0 : (tkn >= p.V ? p.V - 1 : tkn);
        double sum = 0.0;
        for (int f = s; f < e; ++f) {
            const unsigned fl = rec[f];
            const bool gt = fl & 1u, lt = fl & 2u;
            const float *r0 = em_b + (int64_t)f * p.st;
            float pr;
            if (lt && f + 1 < Tb) {  // :106-108 (hard-coded vocabulary index 0, next frame's token)
                const float x = r0[0], y = r0[p.st + tkn];
                pr = expf(fmaxf(x, y));
            } else {
                pr = expf(r0[gt ? tkn : 0]);  // :112
            }
            sum += (double)pr;
            if (ptok) ptok[f] = i;
            if (pprob) pprob[f] = pr;
        }
        en_b[i] = e;
        sc_b[i] = sum / (double)(e - s);
    }
}

static size_t align_smem_bytes(const AlignCfg &c) {
    return align_up(64 + 256 + (size_t)c.stages * c.chunk * c.slot_bytes, 16);
}

}  // namespace ssak

using namespace ssak;

extern "C" size_t ssak_align_workspace_bytes(int64_t B, int64_t Tmax, int64_t Lmax) {
    AlignCfg c;
    if (B <= 0 || Tmax < 0 || Lmax < 0 || !choose_align_cfg(Lmax, B, 64, &c)) return 0;
    return align_up((size_t)B * (size_t)Tmax * 2 * c.NW * sizeof(uint32_t), 256) +
           align_up((size_t)B * (size_t)Tmax, 256) + 256;
}

extern "C" int ssak_forced_align(const float *emissions, int64_t B, int64_t Tmax, int64_t V,
                                 int64_t em_stride_b, int64_t em_stride_t, const int32_t *tokens,
                                 int64_t tok_stride, int64_t Lmax,
                                 const int32_t *emission_lengths, const int32_t *token_lengths,
                                 int32_t blank, int32_t first_as_garbage, const float *col0,
                                 int32_t *starts, int32_t *ends, double *scores, int32_t *t_start,
                                 int32_t *status, float *trellis_dump, int32_t *path_token,
                                 float *path_prob, void *workspace, size_t workspace_bytes,
                                 ssak_stream_t stream) {
    if (!emissions || !tokens || !emission_lengths || !token_lengths || !starts || !ends ||
        !scores || !t_start || !status || !workspace)
        return SSAK_ERR_INVALID_ARGUMENT;
    if (B <= 0 || Tmax < 0 || V <= 0 || Lmax < 0 || blank < 0 || blank >= V ||
        Tmax > 0x7ffffff0 || V > (1 << 20))
        return SSAK_ERR_INVALID_ARGUMENT;
    if (first_as_garbage && !col0) return SSAK_ERR_INVALID_ARGUMENT;
    AlignParams p;
    if (!choose_align_cfg(Lmax, B, (int)V, &p.cfg)) return SSAK_ERR_UNSUPPORTED;
    if (workspace_bytes < ssak_align_workspace_bytes(B, Tmax, Lmax)) return SSAK_ERR_WORKSPACE;
    const size_t smem_bytes = align_smem_bytes(p.cfg);
    if (smem_bytes > 227 * 1024) return SSAK_ERR_UNSUPPORTED;
    p.em = emissions; p.B = B; p.Tmax = Tmax; p.V = (int)V; p.sb = em_stride_b; p.st = em_stride_t;
    p.tokens = tokens; p.tok_stride = tok_stride; p.Lmax = (int)Lmax;
    p.em_len = emission_lengths; p.tok_len = token_lengths; p.blank = blank;
    p.garbage = first_as_garbage; p.col0 = col0;
    char *ws = reinterpret_cast<char *>(workspace);
    p.bp = reinterpret_cast<uint32_t *>(ws);
    p.rec = reinterpret_cast<unsigned char *>(
        ws + align_up((size_t)B * (size_t)Tmax * 2 * p.cfg.NW * sizeof(uint32_t), 256));
    p.starts = starts; p.ends = ends; p.scores = scores; p.t_start = t_start; p.status = status;
    p.dump = trellis_dump; p.path_token = path_token; p.path_prob = path_prob;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)B), block((p.cfg.W + (p.cfg.W < 32 ? 1 : 0)) * 32);
#define SSAK_LAUNCH(KK)                                                                        \
    case KK: {                                                                                 \
        auto kern = align_forward_kernel<KK>;                                                  \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem_bytes);                                 \
        if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }                \
        kern<<<grid, block, smem_bytes, s>>>(p);                                               \
        break;                                                                                 \
    }
    switch (p.cfg.K) {
        SSAK_LAUNCH(1)
        SSAK_LAUNCH(2)
        SSAK_LAUNCH(4)
        SSAK_LAUNCH(8)
        SSAK_LAUNCH(16)
        default: return SSAK_ERR_UNSUPPORTED;
    }
#undef SSAK_LAUNCH
    int rc = check_launch();
    if (rc != SSAK_OK) return rc;
    align_backtrace_kernel<<<(unsigned)B, 128, 0, s>>>(p);
    return check_launch();
}
