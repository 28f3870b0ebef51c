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
    int K, W, NW;  // states per lane, recursion warps, 32-state groups per frame (= K*W)
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
    const float *col0;   // caller's column 0 (first_as_garbage) or nullptr
    float *col0_eff;     // [B][Tmax] column 0 of trellis rows 1..T_b incl. the +inf sentinel (workspace)
    uint32_t *bp;        // [B][Tmax][W][2K] decision bits: per warp K words "changed>stayed", K words "<"
    uint32_t *rec;       // [B][Tmax] (token index << 2) | decision flags of the frames on the path
    float *prob;         // [B][Tmax] per-frame probability of the path (workspace)
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
        while (K < 8 && (P + 32 * K - 1) / (32 * K) > wtarget) K *= 2;
    }
    if (K != 1 && K != 2 && K != 4 && K != 8 && K != 16) return false;
    if (K < 8 && (P + 32 * K - 1) / (32 * K) > 15) K = 8;     // K <= 4 kernels are built for <= 15 recursion warps
    if (K < 16 && (P + 32 * K - 1) / (32 * K) > 31) K = 16;  // one warp of the CTA is the producer
    const int64_t W = (P + 32 * K - 1) / (32 * K);
    if (W > (K == 16 ? 16 : 31)) return false;  // L <= 8191
    c->K = K;
    c->W = (int)W;
    c->NW = K * (int)W;
    c->slot_bytes = ring_slot_bytes(V);
    c->chunk = 8 * c->slot_bytes <= 16384 ? 8 : 4;
    int stages = (64 * 1024) / (c->chunk * c->slot_bytes);
    c->stages = stages > 4 ? 4 : stages;
    return c->stages >= 2;
}

// Column 0 of the trellis (:37 / :39) for rows 1..T_b, with the +inf sentinel of :42.  The cumulative
// variant is the fp64 running sum of the blank column rounded to fp32 per element (torch.cumsum on the
// CPU): strictly sequential by definition, so one warp per utterance fetches 32 frames at a time and
// runs the 32 dependent fp64 adds through shuffles.
__global__ void __launch_bounds__(32) align_col0_kernel(const AlignParams p) {
    const int b = blockIdx.x, lane = threadIdx.x;
    int Tb = p.em_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.Tmax ? (int)p.Tmax : Tb);
    int L = p.tok_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    const float INF = __int_as_float(0x7f800000);
    const float *em_b = p.em + (int64_t)b * p.sb + p.blank;
    float *out = p.col0_eff + (int64_t)b * p.Tmax;
    const float *c0 = p.garbage && p.col0 ? p.col0 + (int64_t)b * p.Tmax : nullptr;
    double acc = 0.0;
    for (int t0 = 0; t0 < Tb; t0 += 32) {
        const int t = t0 + lane;
        float x = 0.f;
        if (t < Tb) x = c0 ? c0[t] : em_b[(int64_t)t * p.st];
        float mine = x;
        if (!c0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                acc += (double)__shfl_sync(0xffffffffu, x, i);
                if (lane == i) mine = (float)acc;
            }
        }
        if (t < Tb) out[t] = (t + 1 >= Tb + 1 - L) ? INF : mine;
    }
}

// Warp roles: [0, W) recursion, W producer (bulk copies of the emission rows, mbarriers only).
template <int K, int CH>
__global__ void __launch_bounds__(K == 16 ? 544 : (K == 8 ? 1024 : 512), 1)
align_forward_kernel(const AlignParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const AlignCfg &c = p.cfg;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);
    const int W = c.W;
    const bool compute = warp < W;

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

    uint64_t *em_full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *em_empty = reinterpret_cast<uint64_t *>(smem + 64);
    float *xchg = reinterpret_cast<float *>(smem + 128);  // [2][34]: guard, W seams
    RowRing ring;
    ring.slots = smem + 416;
    ring.full = em_full;
    ring.chunk = CH;
    ring.stages = c.stages;
    ring.slot_bytes = c.slot_bytes;
    ring.row_bytes = 4 * V;
    const int NST = c.stages, slot_bytes = c.slot_bytes;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&em_full[s], 1);
            mbar_init(&em_empty[s], W);
        }
        mbar_fence_init();
    }
    if (tid < 68) xchg[tid] = -INF;

    // states of this thread, their tokens, trellis row 0 (:35, :41, :42)
    const int jbase = warp * 32 * K + lane;
    int tok_off[K];
    float v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = jbase + k * 32;
        int tkn = p.blank;
        if (compute && j >= 1 && j <= L) {
            tkn = tk[j - 1];
            tkn = tkn < 0 ? 0 : (tkn >= V ? V - 1 : tkn);
        }
        tok_off[k] = 4 * tkn;
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
    __syncthreads();
    if (compute && lane == 31) xchg[1 + warp] = v[K - 1];
    __syncthreads();  // last CTA-wide barrier

    const int nchunks = (Tb + CH - 1) / CH;
    if (!compute) {
        // ================= producer warp =================
        RingProducer prod;
        prod.src = em_b;
        prod.step_elems = p.st;
        prod.stage = 0;
        prod.remaining = Tb;
        int round = 0;
        for (int n = 0; n < nchunks; ++n) {
            if (round > 0) mbar_wait(&em_empty[prod.stage], (uint32_t)((round - 1) & 1));  // hardware-suspended wait
            const int stg = prod.stage;
            if (lane == 0) ring_issue_next(ring, prod);
            prod.stage = __shfl_sync(FULL, prod.stage, 0);
            if (prod.stage <= stg) ++round;
        }
        return;
    }

    // ================= recursion warps: chunk-unrolled time loop =================
    const int nbar = W * 32, NW = c.NW;
    const unsigned a15_0 = (unsigned)(reinterpret_cast<uintptr_t>(em_b) & 15);
    const unsigned a15_step = (unsigned)((p.st * 4) & 15);
    const int blank_off = 4 * p.blank;
    const unsigned seam_m = lane == 0 ? 0xffffffffu : 0u;
    const unsigned zero_m = tid == 0 ? 0xffffffffu : 0u;  // the thread that owns trellis column 0
    auto sel = [](unsigned m, float a, float bb) {
        return __int_as_float((__float_as_int(a) & m) | (__float_as_int(bb) & ~m));
    };
    const float *x_in = xchg + warp;  // seam of warp-1 (index 0 is the guard)
    float *x_out = xchg + 1 + warp;
    uint32_t *bp_ptr = p.bp + (int64_t)b * p.Tmax * 2 * NW + warp * 2 * K;
    const float *c0_ptr = p.col0_eff + (int64_t)b * p.Tmax;
    float *dump_row = p.dump ? p.dump + ((int64_t)b * (p.Tmax + 1) + 1) * (p.Lmax + 1) + jbase : nullptr;
    const unsigned char *em_base = ring.slots, *em_chunk = em_base;
    int em_stage = 0, em_phase = 0, remaining = Tb, t = 0;
    // column-0 values: lane f of warp 0 fetches the value of frame f of the NEXT chunk (one chunk ahead);
    // thread 0 picks its frame's value with a shuffle
    float c0n = (warp == 0 && lane < CH && lane < Tb) ? __ldg(c0_ptr + lane) : 0.f;

    while (remaining > 0) {
        const int n = remaining < CH ? remaining : CH;
        const float c0cur = c0n;
        c0n = (warp == 0 && lane < CH && t + CH + lane < Tb) ? __ldg(c0_ptr + t + CH + lane) : 0.f;
        mbar_wait(&em_full[em_stage], (uint32_t)em_phase);
        constexpr int kUnroll = K >= 8 ? 1 : CH;  // many states per lane: the frame body is big enough
#pragma unroll kUnroll
        for (int f = 0; f < CH; ++f) {
            if (f >= n) break;
            const unsigned char *row = em_chunk + f * slot_bytes + ((a15_0 + f * a15_step) & 15u);
            const float eb = *reinterpret_cast<const float *>(row + blank_off);
            const float xin = x_in[(f & 1) * 34];
            // states in ascending order; the neighbour of state k is the OLD value of the previous state:
            // shuffle v[k] before overwriting it, keep the previous shuffle for the warp seam (lane 0).
            // Decision words are stored as soon as four of a kind are complete (few live registers).
            float rprev = xin;
            uint32_t wg[4], wl[4];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float rk = __shfl_sync(FULL, v[k], (lane + 31) & 31);
                const float prev = sel(seam_m, rprev, rk);
                rprev = rk;
                const float ek = *reinterpret_cast<const float *>(row + tok_off[k]);
                const float stayb = v[k] + eb;             // :48
                const float stayt = v[k] + ek;             // :49
                const float chg = prev + ek;               // :51
                const float stayed = fmaxf(stayb, stayt);  // what backtrack recomputes (:96-99)
                float nv = fmaxf(stayed, chg);
                if (k == 0) nv = sel(zero_m, __shfl_sync(FULL, c0cur, f), nv);  // column 0 (:37 / :39 / :42)
                v[k] = nv;
                wg[k & 3] = __ballot_sync(FULL, chg > stayed);
                wl[k & 3] = __ballot_sync(FULL, chg < stayed);
                if (K == 1) {
                    if (lane == 0) *reinterpret_cast<uint2 *>(bp_ptr) = make_uint2(wg[0], wl[0]);
                } else if (K == 2) {
                    if (k == 1 && lane == 0) *reinterpret_cast<uint4 *>(bp_ptr) = make_uint4(wg[0], wg[1], wl[0], wl[1]);
                } else if ((k & 3) == 3 && lane == 0) {
                    *reinterpret_cast<uint4 *>(bp_ptr + (k - 3)) = make_uint4(wg[0], wg[1], wg[2], wg[3]);
                    *reinterpret_cast<uint4 *>(bp_ptr + K + (k - 3)) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
                }
            }
            if (lane == 31) x_out[((f & 1) ^ 1) * 34] = v[K - 1];
            bp_ptr += 2 * NW;
            if (ownsL) {
                float vl = v[0];
#pragma unroll
                for (int k = 1; k < K; ++k)
                    if (k == kL) vl = v[k];
                if (vl > best) {  // first maximum (:88)
                    best = vl;
                    best_t = t + f + 1;
                }
            }
            if (dump_row) {
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (jbase + k * 32 <= L) dump_row[k * 32] = v[k];
                dump_row += p.Lmax + 1;
            }
            if (f == n - 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&em_empty[em_stage]);
            }
            named_bar_sync(1, nbar);
        }
        remaining -= n;
        t += n;
        em_chunk += CH * slot_bytes;
        if (++em_stage == NST) { em_stage = 0; em_phase ^= 1; em_chunk = em_base; }
    }
    if (ownsL) p.t_start[b] = best_t;
}

// Back-trace (:79-123) + merge_repeats (:141-157).  Warp 0 walks the decision bits from (t_start, L): each
// lane fetches the 64-state window of one frame, 32 frames per round trip, and the fetch for the next 32
// frames is issued before the current 32 are walked (the window [j-63, j] covers wherever the walk ends).
// Then the whole CTA computes the per-frame probabilities (:106-112) in parallel and the per-token means.
__global__ void __launch_bounds__(256) align_backtrace_kernel(const AlignParams p) {
    __shared__ int s_status, s_first;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    int Tb = p.em_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.Tmax ? (int)p.Tmax : Tb);
    int L = p.tok_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    const int NW = p.cfg.NW, KK = p.cfg.K;
    const uint32_t *bp_b = p.bp + (int64_t)b * p.Tmax * 2 * NW;
    uint32_t *rec = p.rec + (int64_t)b * p.Tmax;   // (token index << 2) | (changed>stayed) | (changed<stayed) << 1
    float *prob = p.prob + (int64_t)b * p.Tmax;
    int32_t *st_b = p.starts + (int64_t)b * p.Lmax;
    int32_t *en_b = p.ends + (int64_t)b * p.Lmax;
    double *sc_b = p.scores + (int64_t)b * p.Lmax;
    const int t_start = (L == 0 || Tb == 0) ? 0 : p.t_start[b];

    if (warp == 0) {
        int t = t_start, j = L;
        bool done = false;
        uint32_t cg[3], cl[3], ng[3] = {0, 0, 0}, nl[3] = {0, 0, 0};
        int cgrp = 0, ngrp = 0;
        auto fetch = [&](int tt, int jj, uint32_t *g, uint32_t *l, int &grp0) {
            // words of trellis row tt-lane covering the states [jj-63, jj]; 32-state group g lives at
            // word (g / K) * 2K + (g % K) ("changed > stayed") and + K ("changed < stayed")
            grp0 = max(jj - 63, 0) >> 5;
            const int rr = tt - lane;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                g[q] = 0;
                l[q] = 0;
                const int grp = grp0 + q;
                if (rr >= 1 && grp < NW) {
                    const uint32_t *w = bp_b + (int64_t)(rr - 1) * 2 * NW + (grp / KK) * 2 * KK + (grp % KK);
                    g[q] = __ldg(w);
                    l[q] = __ldg(w + KK);
                }
            }
        };
        if (t > 0) fetch(t, j, cg, cl, cgrp);
        while (t > 0 && !done) {
            if (t > 32) fetch(t - 32, j, ng, nl, ngrp);  // next 32 frames, issued before the walk
            const int base = max(j - 31, 0);
            const int off = base - (cgrp << 5);          // 0 <= off < 64
            const int sh = off & 31;
            const bool hi = off >= 32;
            const uint32_t wg = __funnelshift_r(hi ? cg[1] : cg[0], hi ? cg[2] : cg[1], sh);
            const uint32_t wl = __funnelshift_r(hi ? cl[1] : cl[0], hi ? cl[2] : cl[1], sh);
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
                if (t - i < 1) break;
                const uint32_t gi = __shfl_sync(FULL, wg, i), li = __shfl_sync(FULL, wl, i);
                const int bit = j - base;
                const uint32_t gt = (gi >> bit) & 1u, lt = (li >> bit) & 1u;
                const int frame = t - i - 1;
                if (lane == 0) rec[frame] = ((uint32_t)(j - 1) << 2) | gt | (lt << 1);
                if (gt) {  // :117 changed > stayed -> previous token
                    if (lane == 0) st_b[j - 1] = frame;
                    --j;
                    if (j == 0) { done = true; s_first = frame; break; }  // :119-120
                }
            }
            t -= 32;
#pragma unroll
            for (int q = 0; q < 3; ++q) { cg[q] = ng[q]; cl[q] = nl[q]; }
            cgrp = ngrp;
        }
        if (lane == 0) {
            s_status = done ? 0 : 1;  // :121-122 "Failed to align"
            p.status[b] = done ? 0 : 1;
        }
    }
    __syncthreads();
    const bool ok = s_status == 0;
    const int first = ok ? s_first : 0, last = ok ? t_start : 0;   // path frames: [first, last)
    int32_t *ptok = p.path_token ? p.path_token + (int64_t)b * p.Tmax : nullptr;
    float *pprob = p.path_prob ? p.path_prob + (int64_t)b * p.Tmax : nullptr;
    const float *em_b = p.em + (int64_t)b * p.sb;
    const int32_t *tk = p.tokens + (int64_t)b * p.tok_stride;
    // per-frame probability of the path (:106-112), one thread per frame
    for (int f = tid; f < (int)p.Tmax; f += blockDim.x) {
        float pr = 0.f;
        int ti = -1;
        if (f >= first && f < last) {
            const uint32_t rc = rec[f];
            ti = (int)(rc >> 2);
            int tkn = tk[ti];
            tkn = tkn < 0 ? 0 : (tkn >= p.V ? p.V - 1 : tkn);
            const bool gt = rc & 1u, lt = rc & 2u;
            const float *r0 = em_b + (int64_t)f * p.st;
            if (lt && f + 1 < Tb) {  // hard-coded vocabulary index 0, next frame's token (:108)
                const float x = r0[0], y = r0[p.st + tkn];
                pr = expf(fmaxf(x, y));
            } else {
                pr = expf(r0[gt ? tkn : 0]);  // :112
            }
            prob[f] = pr;
        }
        if (ptok) ptok[f] = ti;
        if (pprob) pprob[f] = pr;
    }
    __syncthreads();
    // merge_repeats (:141-157): span of token i and the mean of its per-frame probabilities
    for (int i = tid; i < p.Lmax; i += blockDim.x) {
        if (!ok || i >= L) {
            st_b[i] = -1;
            en_b[i] = -1;
            sc_b[i] = 0.0;
            continue;
        }
        const int s = st_b[i];
        const int e = i + 1 < L ? st_b[i + 1] : t_start;
        double sum = 0.0;
        for (int f = s; f < e; ++f) sum += (double)prob[f];
        en_b[i] = e;
        sc_b[i] = sum / (double)(e - s);
    }
}

static size_t align_smem_bytes(const AlignCfg &c) {
    return align_up(416 + (size_t)c.stages * c.chunk * c.slot_bytes, 16);
}

}  // namespace ssak

using namespace ssak;

extern "C" size_t ssak_align_workspace_bytes(int64_t B, int64_t Tmax, int64_t Lmax) {
    AlignCfg c;
    if (B <= 0 || Tmax < 0 || Lmax < 0 || !choose_align_cfg(Lmax, B, 64, &c)) return 0;
    return align_up((size_t)B * (size_t)Tmax * 2 * c.NW * sizeof(uint32_t), 256) +
           3 * align_up((size_t)B * (size_t)Tmax * sizeof(float), 256) + 256;
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
    const size_t bp_bytes = align_up((size_t)B * (size_t)Tmax * 2 * p.cfg.NW * sizeof(uint32_t), 256);
    const size_t bt_bytes = align_up((size_t)B * (size_t)Tmax * sizeof(float), 256);
    p.rec = reinterpret_cast<uint32_t *>(ws + bp_bytes);
    p.prob = reinterpret_cast<float *>(ws + bp_bytes + bt_bytes);
    p.col0_eff = reinterpret_cast<float *>(ws + bp_bytes + 2 * bt_bytes);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return SSAK_ERR_INVALID_ARGUMENT;
    p.starts = starts; p.ends = ends; p.scores = scores; p.t_start = t_start; p.status = status;
    p.dump = trellis_dump; p.path_token = path_token; p.path_prob = path_prob;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    align_col0_kernel<<<(unsigned)B, 32, 0, s>>>(p);
    int rc = check_launch();
    if (rc != SSAK_OK) return rc;
    dim3 grid((unsigned)B), block((p.cfg.W + 1) * 32);
#define SSAK_LAUNCH2(KK, CC)                                                                   \
    {                                                                                          \
        auto kern = align_forward_kernel<KK, CC>;                                              \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem_bytes);                                 \
        if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }                \
        kern<<<grid, block, smem_bytes, s>>>(p);                                               \
    }
#define SSAK_LAUNCH(KK)                                                                        \
    case KK:                                                                                   \
        if (p.cfg.chunk == 8) SSAK_LAUNCH2(KK, 8) else SSAK_LAUNCH2(KK, 4)                     \
        break;
    switch (p.cfg.K) {
        SSAK_LAUNCH(1)
        SSAK_LAUNCH(2)
        SSAK_LAUNCH(4)
        SSAK_LAUNCH(8)
        SSAK_LAUNCH(16)
        default: return SSAK_ERR_UNSUPPORTED;
    }
#undef SSAK_LAUNCH
#undef SSAK_LAUNCH2
    rc = check_launch();
    if (rc != SSAK_OK) return rc;
    align_backtrace_kernel<<<(unsigned)B, 256, 0, s>>>(p);
    return check_launch();
}
