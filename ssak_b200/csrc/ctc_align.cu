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
#include <type_traits>

#include "common.cuh"

namespace ssak {

struct AlignCfg {
    int K, W, NW;  // states per lane, recursion warps per CTA, 32-state groups per frame (= K*W*S)
    int chunk, stages, slot_bytes;
    int S;         // CTAs per utterance (wave kernel: one thread-block cluster), 1 for the barrier kernel
    int wave;      // 1: wavefront kernel (warps skewed in time, no per-frame barrier), 0: barrier kernel
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
    float *gseam;         // wave kernel: [B][S-1][Tmax] seam values handed from CTA c to CTA c+1 (preset to NaN)
    int *abort_word;      // wave kernel: kNoAbort until a seam poll gave up (watchdog); then every warp drains and exits
    AlignCfg cfg;
};

// throughput kernel (align_lane_kernel): frames per chunk, chunks of emission rows in shared memory
constexpr int LANE_C = 4, LANE_DEPTH = 4;
__host__ __device__ inline int lane_ring_bytes(int V) { return LANE_DEPTH * LANE_C * (32 * (V <= 64 ? 2 : 4) + 4) * 4; }

static bool choose_barrier_cfg(int64_t Lmax, int64_t B, int V, AlignCfg *c) {
    const int64_t P = Lmax + 1;
    c->S = 1;
    c->wave = 0;
    const int sms = device_sm_count();
    int wtarget = (B <= sms) ? 8 : ((B <= 4 * sms) ? 4 : 2);
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

static inline int align_env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Wavefront kernel shape.  (K, W, S) depend on (Lmax, B) only, so the workspace size does not depend on V;
// V decides the ring geometry and whether the ring fits at all (otherwise the barrier kernel is used).
//   S  CTAs per utterance (a cluster): as many as the machine has room for when there are few utterances,
//      as long as every CTA keeps >= 128 states;
//   W  recursion warps per CTA, K states per lane: 32*K*W*S >= Lmax+1.
static bool choose_wave_shape(int64_t Lmax, int64_t B, AlignCfg *c) {
    const int64_t P = Lmax + 1;
    // up to 16 CTAs per utterance: with B * S <= #SMs the whole grid is resident at one CTA per SM, so a cooperative
    // launch guarantees co-residency without a cluster (clusters stop at 8 CTAs, and a B200 fits only 15 of
    // those at one CTA per SM)
    const int sms = device_sm_count();
    int S = (int)(sms / (B < 1 ? 1 : B));
    S = S < 1 ? 1 : (S > 16 ? 16 : S);
    while (S > 1 && (P + S - 1) / S < 128) --S;
    const int64_t cap = 32 * 8 * 8;  // states one CTA can hold (K = 8, W = 8)
    if ((P + cap - 1) / cap > S) S = (int)((P + cap - 1) / cap);
    S = align_env_int("SSAK_ALIGN_S", S);
    if (S < 1 || S > 16) return false;
    const int64_t Pc = (P + S - 1) / S;
    int wtarget = (B * S <= sms) ? 8 : ((B * S <= 4 * sms) ? 4 : 2);
    wtarget = align_env_int("SSAK_ALIGN_WARPS", wtarget);
    int K = align_env_int("SSAK_ALIGN_K", 0);
    if (K == 0) {
        // few CTAs (latency regime): a lone warp issues ~0.3 instructions per cycle whatever its ILP, so thinner
        // warps help -- down to 2 states per lane (measured on the C2-shaped batch: 0.26 ms at K = 2, 0.29 ms at
        // K = 1 (longer warp chain, two more shuffles per frame) and at K = 4); many CTAs: 4 states per lane
        // (fewest instructions per state), 8 beyond 1024 states per CTA
        if (B * S <= sms) {
            K = 2;
            while (K < 8 && (Pc + 32 * K - 1) / (32 * K) > 8) K *= 2;
        } else {
            K = (Pc + 127) / 128 > wtarget ? 8 : 4;
        }
    }
    if (K != 1 && K != 2 && K != 4 && K != 8) return false;  // consecutive states per lane
    while (K < 8 && (Pc + 32 * K - 1) / (32 * K) > 8) K *= 2;  // the kernels are built for <= 8 recursion warps
    int64_t W = (Pc + 32 * K - 1) / (32 * K);
    if (W > 8) return false;
    c->K = K;
    c->W = (int)W;
    c->S = S;
    c->NW = K * (int)W * S;
    c->wave = 1;
    return true;
}

static bool choose_wave_ring(int64_t B, int V, AlignCfg *c) {
    c->slot_bytes = ring_slot_bytes(V);
    c->chunk = 8 * c->slot_bytes <= 16384 ? 8 : 4;
    const int budget = (B * c->S <= device_sm_count()) ? 160 * 1024 : 72 * 1024;
    int stages = budget / (c->chunk * c->slot_bytes);
    // warp w works >= w chunks behind warp 0: W+2 stages at least, W+6 cover the bulk-copy latency as well
    if (stages > c->W + 6) stages = c->W + 6;
    stages = align_env_int("SSAK_ALIGN_STAGES", stages);
    if (stages > 24) stages = 24;
    c->stages = stages;
    return stages >= c->W + 2;
}

// Throughput shape (align_lane_kernel): the decision bits are laid out as align_wave_kernel<8> with W = KL / 8
// "virtual" warps, so the back-trace kernel and the workspace layout need nothing new.  V < 0: not known
// (workspace query).  SSAK_ALIGN_LANE=1 forces it wherever it is valid, =0 disables it.
static bool choose_lane_cfg(int64_t Lmax, int64_t B, int V, AlignCfg *c) {
    const int mode = align_env_int("SSAK_ALIGN_LANE", -1);
    if (mode == 0 || Lmax + 1 > 512 || V > 128 || B > 0x7fffffff) return false;
    if (mode < 0 && B < 2 * (int64_t)device_sm_count()) return false;
    c->K = 8;
    c->W = Lmax + 1 <= 256 ? 1 : 2;
    c->S = 1;
    c->NW = 8 * c->W;
    c->wave = 2;
    c->chunk = LANE_C;
    c->stages = LANE_DEPTH;
    c->slot_bytes = 0;
    return true;
}

static bool choose_align_cfg(int64_t Lmax, int64_t B, int V, AlignCfg *c, bool allow_lane = true) {
    if (allow_lane && choose_lane_cfg(Lmax, B, V, c)) return true;
    if (align_env_int("SSAK_ALIGN_WAVE", 1) != 0 && choose_wave_shape(Lmax, B, c) && choose_wave_ring(B, V, c))
        return true;
    return choose_barrier_cfg(Lmax, B, V, c);
}

// bytes of the decision-bit matrix for the larger of the two candidate shapes (the choice depends on V)
static int align_max_nw(int64_t Lmax, int64_t B, int *S_out) {
    AlignCfg a, w;
    int nw = 0;
    *S_out = 1;
    if (choose_barrier_cfg(Lmax, B, 64, &a)) nw = a.NW;
    if (choose_wave_shape(Lmax, B, &w)) {
        nw = w.NW > nw ? w.NW : nw;
        *S_out = w.S;
    }
    if (choose_lane_cfg(Lmax, B, 1, &a)) nw = a.NW > nw ? a.NW : nw;
    return nw;
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
    auto load = [&](int t) -> float { return t < Tb ? (c0 ? c0[t] : em_b[(int64_t)t * p.st]) : 0.f; };
    float xn = load(lane);
    for (int t0 = 0; t0 < Tb; t0 += 32) {
        const int t = t0 + lane;
        const float x = xn;
        xn = load(t + 32);  // the next 32 frames are in flight while this block's chain runs
        float mine = x;
        if (!c0) {
            // the only serial piece is the chain of fp64 adds: convert once per lane, broadcast the doubles
            // (independent shuffles), pick my prefix with selects, round to fp32 once at the end
            const double xd = (double)x;
            double mine_d = 0.0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                acc += __shfl_sync(0xffffffffu, xd, i);
                mine_d = lane == i ? acc : mine_d;
            }
            mine = (float)mine_d;
        }
        if (t < Tb) out[t] = (t + 1 >= Tb + 1 - L) ? INF : mine + 0.0f;  // (+0.0f: a -0.0 becomes +0.0)
    }
}

// Warp roles: [0, W) recursion, W producer (bulk copies of the emission rows, mbarriers only).
template <int K, int CH>
__global__ void __launch_bounds__(K == 16 ? 544 : (K == 8 ? 1024 : 512), 1)
align_forward_kernel(const AlignParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
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

// ---------------------------------------------------------------------------------------------
// Wavefront forward kernel.  Same arithmetic and the same decision-bit layout as the barrier kernel, but
// the recursion warps are skewed in time instead of meeting at a barrier every frame: state j of row t+1
// depends on states j and j-1 of row t only, so warp g (states [32K g, 32K (g+1))) needs exactly ONE value
// per frame from warp g-1 (its "seam").  Warp g-1 streams its seam values into a small shared-memory ring and
// warp g runs a chunk or two behind it.  The chain continues across the CTAs of a thread-block cluster
// (S CTAs per utterance, co-scheduled by the cluster launch) through a global, L2-resident seam buffer.
//   * Seam words validate themselves ("flag in the data"): every word starts as a NaN pattern that no
//     trellis value can take, the consumer polls the word itself.  No flag, no fence, no barrier; each
//     4-byte store is atomic.  The consumer reads the words of the NEXT chunk while it computes the current
//     one and only polls when that early read came too soon.
//   * The emission ring's full-barrier is probed early in the same way (test_wait a chunk ahead, blocking
//     try_wait only if that probe failed), so in the steady state a chunk costs no synchronisation latency.
//   * Nothing ever waits on a downstream warp, so the chain cannot deadlock.  Every CTA has its own emission
//     ring (producer warp W); a stage is recycled once all live warps of the CTA released it, which bounds
//     the lead of warp g-1 over warp g (-> a seam ring of stages+1 chunks never overflows).
//   * The frame body is branch-free (selects and predicated stores only) and unrolled over the chunk, so
//     that ptxas overlaps the loads, ballots and stores of neighbouring frames with the dependent chain
//     shuffle -> add -> max, which is all that remains serial per frame.
struct WaveSmem {
    int em_full, em_empty, seam_val, ring, total;
};
__host__ __device__ __forceinline__ WaveSmem wave_smem(int W, int stages, int chunk, int slot_bytes) {
    WaveSmem m;
    const int nslot = stages + 1;
    m.em_full = 0;
    m.em_empty = 8 * stages;
    m.seam_val = 16 * stages;                      // [W+2][nslot][chunk] floats (+ scratch ring, + column-0 ring)
    m.ring = (m.seam_val + 4 * (W + 2) * nslot * chunk + 127) & ~127;
    m.total = m.ring + stages * chunk * slot_bytes;
    return m;
}

constexpr uint32_t kSeamEmpty = 0xffffffffu;  // NaN pattern no add/max of the recursion produces
__device__ __forceinline__ float ld_seam_global(const float *p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_seam_shared(const float *p) {
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
// Watchdog of the seam polls.  With the co-residency the launch guarantees (cooperative or cluster launch) an
// upstream warp always makes progress, so this never fires; if it does (10 s of WALL time without the upstream
// value, %globaltimer) the poller sets the call's abort word, every poller that sees the word gives up as well,
// the warps drain their emission ring so that the producers finish, and the back-trace kernel reports status 2
// for the utterances of the call -- a status, not a __trap() that would poison the caller's CUDA context.
constexpr int kNoAbort = -1;   // (the abort word is preset by the same 0xff memset as the global seam buffer)
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct SpinGuard {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    // call after every failed poll; true = give up
    __device__ __forceinline__ bool expired(int *abort_word) {
        if ((++spins & 0x3fffu) != 0) return false;
        int a;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(a) : "l"(abort_word) : "memory");
        if (a != kNoAbort) return true;
        const unsigned long long now = global_timer_ns();
        if (t0 == 0) {
            t0 = now;
        } else if (now - t0 > 10000000000ull) {
            atomicExch(abort_word, 1);
            return true;
        }
        return false;
    }
};
__device__ __forceinline__ void st_seam_generic(float *p, float v) {  // shared or global (generic address)
    // no "memory" clobber: nothing in this thread reads the word back, and the clobber would pin every
    // emission load of the unrolled chunk behind the store of the previous frame
    asm volatile("st.relaxed.gpu.f32 [%0], %1;" ::"l"(p), "f"(v));
}

template <int K, int CH, bool DUMP>
__global__ void __launch_bounds__(320, 1) align_wave_kernel(const AlignParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const AlignCfg &c = p.cfg;
    const int cta = blockIdx.x, S = c.S, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);
    const float EMPTY = __uint_as_float(kSeamEmpty);
    const int W = c.W;

    int Tb = p.em_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.Tmax ? (int)p.Tmax : Tb);
    int L = p.tok_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    if (L == 0 || Tb == 0) {  // reference: empty back-track loop -> "Failed to align"
        if (cta != 0) return;
        if (tid == 0) p.t_start[b] = 0;
        if (DUMP) {  // :41-42 with an empty token list / no frames: column 0 is all +inf
            float *d = p.dump + (int64_t)b * (p.Tmax + 1) * (p.Lmax + 1);
            for (int t = tid; t <= Tb; t += blockDim.x) d[(int64_t)t * (p.Lmax + 1)] = INF;
            if (Tb == 0)
                for (int j = 1 + tid; j <= L; j += blockDim.x) d[j] = -INF;
        }
        return;
    }
    // live warps of this CTA: global warp g holds states [32K g, 32K (g+1)), live iff 32K g <= L
    const int gw0 = cta * W;
    int wlive = L / (32 * K) + 1 - gw0;
    wlive = wlive < 0 ? 0 : (wlive > W ? W : wlive);
    if (wlive == 0) return;                        // the whole CTA lies beyond state L
    const bool compute = warp < wlive;

    const int V = p.V;
    const float *em_b = p.em + (int64_t)b * p.sb;
    const int32_t *tk = p.tokens + (int64_t)b * p.tok_stride;
    const int NST = c.stages, NSLOT = NST + 1, slot_bytes = c.slot_bytes;
    const WaveSmem lay = wave_smem(W, NST, CH, slot_bytes);
    uint64_t *em_full = reinterpret_cast<uint64_t *>(smem + lay.em_full);
    uint64_t *em_empty = reinterpret_cast<uint64_t *>(smem + lay.em_empty);
    float *seam_val = reinterpret_cast<float *>(smem + lay.seam_val);          // [W+2][NSLOT][CH]
    RowRing ring;
    ring.slots = smem + lay.ring;
    ring.full = em_full;
    ring.chunk = CH;
    ring.stages = NST;
    ring.slot_bytes = slot_bytes;
    ring.row_bytes = 4 * V;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&em_full[s], 1);
            mbar_init(&em_empty[s], wlive);
        }
        mbar_fence_init();
    }
    for (int i = tid; i < (W + 2) * NSLOT * CH; i += blockDim.x) seam_val[i] = EMPTY;
    __syncthreads();  // the only CTA-wide barrier
    if (warp >= wlive && warp != W && !(warp == W + 1 && cta == 0)) return;  // idle recursion warps

    if (warp == W + 1) {
        // ================= column-0 warp (first CTA of the utterance) =================
        // Column 0 of the trellis (:37 / :39 / :42) is the "seam" of global warp 0: the fp64 running sum of the
        // blank column rounded to fp32 per element (torch.cumsum on the CPU; strictly sequential, 32 frames per
        // round: doubles broadcast by independent shuffles, only the chain of adds is serial), or the caller's
        // vector (first_as_garbage), with the +inf sentinel.  It is streamed into a seam ring like any other
        // seam; a slot is written once its consumer has recycled it.
        float *c0ring = seam_val + (W + 1) * NSLOT * CH;
        const float *c0 = p.garbage && p.col0 ? p.col0 + (int64_t)b * p.Tmax : nullptr;
        const float *blank_col = em_b + p.blank;
        auto load = [&](int tt) -> float { return tt < Tb ? (c0 ? c0[tt] : blank_col[(int64_t)tt * p.st]) : 0.f; };
        double acc = 0.0;
        float xn = load(lane);
        int slot = 0;
        for (int t0 = 0; t0 < Tb; t0 += 32) {
            const int tt = t0 + lane;
            const float x = xn;
            xn = load(tt + 32);
            float mine = x;
            if (!c0) {
                const double xd = (double)x;
                double mine_d = 0.0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    acc += __shfl_sync(FULL, xd, i);
                    mine_d = lane == i ? acc : mine_d;
                }
                mine = (float)mine_d;
            }
            const float val = (tt + 1 >= Tb + 1 - L) ? INF : mine + 0.0f;  // (+0.0f: a -0.0 becomes +0.0)
            for (int j = 0; j < 32 / CH && t0 + j * CH < Tb; ++j) {
                float *dst = c0ring + slot * CH;
                int gave_up = 0;
                if (lane == 0) {
                    SpinGuard guard;
                    while (__float_as_uint(ld_seam_shared(dst)) != kSeamEmpty) {
                        __nanosleep(100);
                        if (guard.expired(p.abort_word)) { gave_up = 1; break; }   // the consumer never recycled the slot
                    }
                }
                if (__shfl_sync(FULL, gave_up, 0)) return;
                if (lane / CH == j && tt < Tb) dst[lane % CH] = val;
                if (++slot == NSLOT) slot = 0;
            }
        }
        return;
    }

    const int nchunks = (Tb + CH - 1) / CH;
    const bool contig = p.st == V;  // [.., T, V] rows back to back: chunk-sized copies, rows 4V bytes apart in the ring
    if (!compute) {
        // ================= producer warp =================
        RingProducer prod;
        prod.src = em_b;
        prod.step_elems = p.st;
        prod.stage = 0;
        prod.remaining = Tb;
        int round = 0;
        for (int n = 0; n < nchunks; ++n) {
            if (round > 0) mbar_wait(&em_empty[prod.stage], (uint32_t)((round - 1) & 1));
            const int stg = prod.stage;
            if (contig) {
                // rows are back to back in memory: the whole chunk is ONE bulk copy (its enclosing 16-byte range)
                if (lane == 0) {
                    const int nf = prod.remaining < CH ? prod.remaining : CH;
                    const uintptr_t a = reinterpret_cast<uintptr_t>(prod.src), a0 = a & ~(uintptr_t)15;
                    const uint32_t bytes = (uint32_t)(((a + (size_t)nf * 4 * V + 15) & ~(uintptr_t)15) - a0);
                    mbar_arrive_expect_tx(&em_full[stg], bytes);
                    bulk_g2s(ring.slots + (size_t)stg * CH * slot_bytes, reinterpret_cast<const void *>(a0), bytes,
                             &em_full[stg]);
                }
                prod.src += (int64_t)CH * V;
                prod.remaining -= CH;
                if (++prod.stage == NST) prod.stage = 0;
            } else {
                if (lane == 0) ring_issue_next(ring, prod);
                prod.stage = __shfl_sync(FULL, prod.stage, 0);
            }
            if (prod.stage <= stg) ++round;
        }
        return;
    }

    // ================= recursion warps =================
    // Blocked state layout: lane l of global warp g holds the K CONSECUTIVE states 32K g + K l + [0, K), so
    // the neighbour of every state but the lane's first one is a register, and ONE shuffle per frame brings
    // the last state of lane l-1 (or the seam value, for lane 0).
    const int gw = gw0 + warp;
    const int sbase = gw * 32 * K + lane * K;
    int tok_off[K];
    float v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = sbase + k;
        int tkn = p.blank;
        if (j >= 1 && j <= L) {
            tkn = tk[j - 1];
            tkn = tkn < 0 ? 0 : (tkn >= V ? V - 1 : tkn);
        }
        tok_off[k] = 4 * tkn;
        v[k] = j == 0 ? (L >= Tb + 1 ? INF : 0.f) : -INF;  // trellis row 0 (:35, :41, :42)
    }
    const int jL_rel = L - gw * 32 * K;  // state L inside this warp?
    const bool ownsL = jL_rel >= 0 && jL_rel < 32 * K && jL_rel / K == lane;
    const int kL = jL_rel & (K - 1);
    float best = -INF;  // trellis[0, L] with L >= 1
    int best_t = 0;
    float *dump_row = nullptr;
    if (DUMP) {
        float *d = p.dump + (int64_t)b * (p.Tmax + 1) * (p.Lmax + 1);
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (sbase + k <= L) d[sbase + k] = v[k];
        dump_row = p.dump + ((int64_t)b * (p.Tmax + 1) + 1) * (p.Lmax + 1) + sbase;
    }
    // seams: where my incoming values come from, where my outgoing values go
    const bool first = gw == 0;                       // owns trellis column 0: its "seam" is column 0 itself
    const bool up_smem = warp > 0 || first, up_glob = warp == 0 && cta > 0;
    const bool dn_live = (gw + 1) * 32 * K <= L;
    const bool dn_glob = dn_live && warp + 1 == W;
    const float *gs_in = p.gseam + ((int64_t)b * (S - 1) + (cta - 1)) * p.Tmax;   // valid when up_glob
    const float *sv_in = seam_val + (first ? W + 1 : warp - 1) * NSLOT * CH;     // valid when up_smem
    // outgoing: the ring of my downstream warp in this CTA, the global buffer, or a scratch ring nobody reads
    float *sv_out = seam_val + (dn_live && !dn_glob ? warp : W) * NSLOT * CH;
    float *out_ptr = dn_glob ? p.gseam + ((int64_t)b * (S - 1) + cta) * p.Tmax : sv_out;
    const bool is31 = lane == 31;

    // frame f of a chunk sits at f * row_stride + ((a15 + f * a15_step) & 15) in its stage: per-row copies land
    // every row (addr & 15) bytes into its own slot, a chunk copy lands the chunk (addr & 15) bytes into the stage
    const int row_stride = contig ? 4 * V : slot_bytes;
    const unsigned a15_step = contig ? 0u : (unsigned)((p.st * 4) & 15);
    const unsigned a15_chunk_step = contig ? (unsigned)((CH * 4 * V) & 15) : 0u;
    unsigned a15 = (unsigned)(reinterpret_cast<uintptr_t>(em_b) & 15);
    const int blank_off = 4 * p.blank;
    const unsigned seam_m = lane == 0 ? 0xffffffffu : 0u;
    const unsigned zero_m = (first && lane == 0) ? 0xffffffffu : 0u;  // the thread that owns trellis column 0
    auto sel = [](unsigned m, float a, float bb) {
        return __int_as_float((__float_as_int(a) & m) | (__float_as_int(bb) & ~m));
    };
    // decision bits, "lane entry" layout: per frame and lane one entry of 2K bits -- bit k: changed > stayed for
    // my k-th state, bit K+k: changed < stayed -- i.e. one byte (K = 4) or two (K = 8) per lane, a coalesced
    // 32/64-byte store per warp and frame, 16 states per 32-bit word, no cross-lane packing at all.
    // K = 1, 2: the 4/K lanes that share a byte OR their bits together (1-2 xor shuffles) and the first one stores,
    // so the layout is that of K = 4 (state s -> byte s/4, "greater" bit s%4, "less" bit 4 + s%4).
    using bp_t = typename std::conditional<K == 8, uint16_t, uint8_t>::type;
    constexpr int LPB = K < 4 ? 4 / K : 1;                  // lanes per entry
    const int64_t bp_row = (int64_t)c.S * W * (32 / LPB);  // entries per frame
    bp_t *bp_ptr = reinterpret_cast<bp_t *>(p.bp) + (int64_t)b * p.Tmax * bp_row + gw * (32 / LPB) + lane / LPB;
    const int ent_sh = K * (lane % LPB);                    // my bits' position inside the nibble
    const bool ent_store = lane % LPB == 0;
    const unsigned char *em_base = ring.slots, *em_chunk = em_base;
    int em_stage = 0, em_phase = 0, remaining = Tb, t = 0, sslot = 0;
    const bool inlane = lane < CH;
    // incoming values of the NEXT chunk, frame f in lane f (read one chunk early; EMPTY = not there yet)
    auto fetch_seam = [&](int t_next, int slot_next) -> float {
        float x = -INF;
        if (up_smem) {
            if (inlane) x = ld_seam_shared(sv_in + slot_next * CH + lane);
        } else if (up_glob) {
            if (inlane && t_next + lane < Tb) x = ld_seam_global(gs_in + t_next + lane);
        }
        return x;
    };
    float sv_pre = fetch_seam(0, 0);

    // one frame.  xin: lane 0's neighbour (state 32K g - 1 in row t+f), or column 0 for the first warp.
    // My last state's value in row t+f is the downstream warp's neighbour at this frame.  A store inside the
    // unrolled chunk would pin every emission load behind it (a generic / shared store may alias the ring as
    // far as ptxas can tell), so full chunks collect the values in registers and store them after the chunk.
    float sout[CH];
    auto frame = [&](const int f, const float sv, auto direct_tag) {
        constexpr bool DIRECT = decltype(direct_tag)::value;
        const unsigned char *row = em_chunk + f * row_stride + ((a15 + f * a15_step) & 15u);
        const float eb = *reinterpret_cast<const float *>(row + blank_off);
        const float xin = __shfl_sync(FULL, sv, f);
        if (DIRECT) {
            if (is31) st_seam_generic(out_ptr + f, v[K - 1]);
        } else {
            sout[f] = v[K - 1];
        }
        const float rk = __shfl_sync(FULL, v[K - 1], (lane + 31) & 31);
        float prev = sel(seam_m, xin, rk);
        // Decision bits without predicates (FSETP -> SEL pairs serialise on the 7 predicate registers): the sign of
        // chg - stayed is "changed < stayed", the sign of stayed - chg is "changed > stayed"; a tie gives +0 both
        // ways and (-inf) - (-inf) gives the positive canonical NaN, i.e. no bit, exactly like the comparisons.
        // One funnel shift pushes a sign bit into the entry.  (A -0.0 could only enter through a caller-supplied
        // column 0; the col0 kernel turns it into +0.0.)
        unsigned ng = 0, nl = 0;
        float stayed_k[K], chg_k[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float ek = *reinterpret_cast<const float *>(row + tok_off[k]);
            // max(v+eb, v+ek) == v + max(eb, ek) bit for bit (rounding is monotonic): :48-49, :96-99
            const float stayed = v[k] + fmaxf(eb, ek);
            const float chg = prev + ek;               // :51
            prev = v[k];
            float nv = fmaxf(stayed, chg);
            if (k == 0) nv = sel(zero_m, xin, nv);     // column 0 (:37 / :39 / :42)
            v[k] = nv;
            stayed_k[k] = stayed;
            chg_k[k] = chg;
        }
#pragma unroll
        for (int k = K - 1; k >= 0; --k) {
            ng = __funnelshift_l(__float_as_uint(stayed_k[k] - chg_k[k]), ng, 1);
            nl = __funnelshift_l(__float_as_uint(chg_k[k] - stayed_k[k]), nl, 1);
        }
        if (K >= 4) {
            *bp_ptr = (bp_t)(ng | (nl << K));
        } else {
            unsigned e = (ng << ent_sh) | (nl << (4 + ent_sh));
            e |= __shfl_xor_sync(FULL, e, 1);
            if (K == 1) e |= __shfl_xor_sync(FULL, e, 2);
            if (ent_store) *bp_ptr = (bp_t)e;
        }
        bp_ptr += bp_row;
        {   // first maximum of trellis[:, L] (:88); only the owner's comparison can be true
            float vl = v[0];
#pragma unroll
            for (int k = 1; k < K; ++k) vl = k == kL ? v[k] : vl;
            const bool up = ownsL && vl > best;
            best = up ? vl : best;
            best_t = up ? t + f + 1 : best_t;
        }
        if (DUMP) {
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (sbase + k <= L) dump_row[k] = v[k];
            dump_row += p.Lmax + 1;
        }
    };

    while (remaining > 0) {
        const int n = remaining < CH ? remaining : CH;
        // incoming values of this chunk: read early during the previous chunk, poll only if that was too soon
        float sv = sv_pre;
        {
            SpinGuard guard;
            bool gave_up = false;
            while (__any_sync(FULL, lane < n && __float_as_uint(sv) == kSeamEmpty)) {
                sv = up_smem ? ld_seam_shared(sv_in + sslot * CH + (lane & (CH - 1)))
                             : ld_seam_global(gs_in + min(t + (lane & (CH - 1)), Tb - 1));
                if (__any_sync(FULL, guard.expired(p.abort_word))) { gave_up = true; break; }
            }
            if (gave_up) {
                // watchdog: drain the emission ring (the producer waits for every live warp) and leave
                while (remaining > 0) {
                    mbar_wait(&em_full[em_stage], (uint32_t)em_phase);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&em_empty[em_stage]);
                    remaining -= remaining < CH ? remaining : CH;
                    if (++em_stage == NST) { em_stage = 0; em_phase ^= 1; }
                }
                return;
            }
            if (up_smem && inlane) const_cast<float *>(sv_in)[sslot * CH + lane] = EMPTY;  // recycle the slot
        }
        const int sslot_next = sslot + 1 == NSLOT ? 0 : sslot + 1;
        sv_pre = fetch_seam(t + CH, sslot_next);
        mbar_wait(&em_full[em_stage], (uint32_t)em_phase);
        if (dn_live && !dn_glob) out_ptr = sv_out + sslot * CH;
        if (n == CH) {
#pragma unroll
            for (int f = 0; f < CH; ++f) frame(f, sv, std::false_type{});
            if (is31) {
#pragma unroll
                for (int f = 0; f < CH; ++f) st_seam_generic(out_ptr + f, sout[f]);
            }
        } else {
#pragma unroll 1
            for (int f = 0; f < n; ++f) frame(f, sv, std::true_type{});
        }
        if (dn_glob) out_ptr += CH;
        a15 = (a15 + a15_chunk_step) & 15u;
        __syncwarp();
        if (lane == 0) mbar_arrive(&em_empty[em_stage]);
        remaining -= n;
        t += n;
        em_chunk += CH * slot_bytes;
        if (++em_stage == NST) { em_stage = 0; em_phase ^= 1; em_chunk = em_base; }
        sslot = sslot_next;
    }
    if (ownsL) p.t_start[b] = best_t;
}

// ------------------------------------------------------------------------------ throughput forward kernel
// Batches that fill the GPU (B >= 2 x SMs) with targets up to 511 tokens and V <= 128: ONE WARP per utterance,
// one CTA per warp (the block scheduler hands an SM a new utterance the moment one finishes).  Lane l holds the KL
// (8 or 16) CONSECUTIVE trellis states [KL l, KL l + KL) in registers, a frame needs one shuffle (the last state of
// lane l-1) and nothing else: no seams, no polling, no mbarriers, no producer warp.  The warp stages its own emission
// rows with cp.async (no registers, no scoreboard; completion counted per commit group) into a ring of
// LANE_DEPTH chunks of 4 frames and computes column 0 itself, 32 rows at a time.  Same arithmetic and the
// same decision-bit layout as align_wave_kernel<8> (lane entries of 16 bits, 16 states per 32-bit word: a lane of
// 16 states writes one word per frame, a coalesced 128-byte row), so align_backtrace_kernel<8> reads it unchanged.

template <int KL, int NV>
__global__ void __launch_bounds__(32, 16) align_lane_kernel(const AlignParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int RS = 32 * NV + 4, C = LANE_C, DEPTH = LANE_DEPTH;
    const int b = blockIdx.x, lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);
    int Tb = p.em_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.Tmax ? (int)p.Tmax : Tb);
    int L = p.tok_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    if (L == 0 || Tb == 0) {  // reference: empty back-track loop -> "Failed to align"
        if (lane == 0) p.t_start[b] = 0;
        return;
    }
    const int V = p.V;
    const float *em_b = p.em + (int64_t)b * p.sb;
    const int32_t *tk = p.tokens + (int64_t)b * p.tok_stride;
    float *ring = reinterpret_cast<float *>(smem);
    const int sbase = lane * KL;
    int tok_off[KL];
    float v[KL];
    unsigned ownL[KL];       // all ones for the one register that holds state L
#pragma unroll
    for (int k = 0; k < KL; ++k) {
        const int j = sbase + k;
        int tkn = p.blank;
        if (j >= 1 && j <= L) {
            tkn = tk[j - 1];
            tkn = tkn < 0 ? 0 : (tkn >= V ? V - 1 : tkn);
        }
        tok_off[k] = 4 * tkn;
        v[k] = j == 0 ? (L >= Tb + 1 ? INF : 0.f) : -INF;  // trellis row 0 (:35, :41, :42)
        ownL[k] = j == L ? 0xffffffffu : 0u;
        // (opaque: otherwise the compiler re-derives the masks as predicates, an ISETP + SEL pair per state and frame)
        asm volatile("" : "+r"(ownL[k]));
    }
    const bool ownsL = L >= sbase && L < sbase + KL;
    float best = -INF;  // trellis[0, L] with L >= 1
    int best_t = 0;
    const int blank_off = 4 * p.blank;
    const unsigned zero_m = lane == 0 ? 0xffffffffu : 0u;  // the thread that owns trellis column 0
    auto sel = [](unsigned m, float a, float bb) {
        return __int_as_float((__float_as_int(a) & m) | (__float_as_int(bb) & ~m));
    };
    // decision bits: 16-bit entries (bit k: changed > stayed of the entry's k-th state, bit 8+k: changed < stayed)
    using bp_t = typename std::conditional<KL == 16, uint32_t, uint16_t>::type;
    constexpr int64_t bp_row = 32;                          // stores per frame (one per lane)
    bp_t *bp_ptr = reinterpret_cast<bp_t *>(p.bp) + (int64_t)b * p.Tmax * bp_row + lane;

    auto issue = [&](int n) {                               // chunk n -> ring stage n % DEPTH (empty group beyond the end)
        const int t0 = n * C;
        const float *src = em_b + (int64_t)t0 * p.st + lane;
        float *dst = ring + (size_t)(n % DEPTH) * C * RS + lane;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            if (t0 + i < Tb) {
#pragma unroll
                for (int jv = 0; jv < NV; ++jv)
                    if (lane + 32 * jv < V)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + i * RS + 32 * jv)),
                                     "l"(src + 32 * jv) : "memory");
            }
            src += p.st;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int nchunks = (Tb + C - 1) / C;
#pragma unroll
    for (int n = 0; n < DEPTH - 1; ++n) issue(n);
    // Column 0 of the trellis (:37 / :39 / :42), 32 rows at a time, row 32 r + i + 1 in lane i: the fp64 running sum
    // of the blank column rounded to fp32 per element (torch.cumsum on the CPU) -- strictly sequential, so the doubles
    // are broadcast by independent shuffles and only the chain of adds is serial -- or the caller's vector
    // (first_as_garbage), with the +inf sentinel.  The next block's inputs are fetched a block ahead.
    const float *c0src = p.garbage && p.col0 ? p.col0 + (int64_t)b * p.Tmax : nullptr;
    const float *blank_col = em_b + p.blank;
    auto c0_load = [&](int tt) -> float { return tt < Tb ? (c0src ? __ldg(c0src + tt) : __ldg(blank_col + (int64_t)tt * p.st)) : 0.f; };
    double c0_acc = 0.0;
    float c0_in = c0_load(lane), c0 = 0.f;
    for (int n = 0; n < nchunks; ++n) {
        const int t0 = n * C, nf = Tb - t0 < C ? Tb - t0 : C;
        __syncwarp();                                       // every lane is done with the previous chunk's rows
        issue(n + DEPTH - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        __syncwarp();
        if ((t0 & 31) == 0) {
            const float x = c0_in;
            c0_in = c0_load(t0 + 32 + lane);
            float mine = x;
            if (!c0src) {
                const double xd = (double)x;
                double mine_d = 0.0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    c0_acc += __shfl_sync(FULL, xd, i);
                    mine_d = lane == i ? c0_acc : mine_d;
                }
                mine = (float)mine_d;
            }
            c0 = (t0 + lane + 1 >= Tb + 1 - L) ? INF : mine + 0.0f;  // (+0.0f: a -0.0 becomes +0.0)
        }
        const unsigned char *rows = reinterpret_cast<const unsigned char *>(ring + (size_t)(n % DEPTH) * C * RS);
        // (not unrolled over the frames: the body is ~10 instructions x KL states already)
#pragma unroll 1
        for (int f = 0; f < nf; ++f) {
            const unsigned char *row = rows + f * (RS * 4);
            const float eb = *reinterpret_cast<const float *>(row + blank_off);
            const float xin = __shfl_sync(FULL, c0, (t0 + f) & 31);   // column 0 of this row
            float prev = __shfl_up_sync(FULL, v[KL - 1], 1);
            unsigned ng = 0, nl = 0, ng2 = 0, nl2 = 0;
            float stayed_k[KL], chg_k[KL];
#pragma unroll
            for (int k = 0; k < KL; ++k) {
                const float ek = *reinterpret_cast<const float *>(row + tok_off[k]);
                // max(v+eb, v+ek) == v + max(eb, ek) bit for bit (rounding is monotonic): :48-49, :96-99
                const float stayed = v[k] + fmaxf(eb, ek);
                const float chg = prev + ek;               // :51
                prev = v[k];
                float nv = fmaxf(stayed, chg);
                if (k == 0) nv = sel(zero_m, xin, nv);
                v[k] = nv;
                stayed_k[k] = stayed;
                chg_k[k] = chg;
            }
            // sign bits as decision bits (see align_wave_kernel): a tie gives +0 both ways, (-inf) - (-inf) the positive
            // canonical NaN, i.e. no bit, exactly like the comparisons
#pragma unroll
            for (int k = 7; k >= 0; --k) {
                ng = __funnelshift_l(__float_as_uint(stayed_k[k] - chg_k[k]), ng, 1);
                nl = __funnelshift_l(__float_as_uint(chg_k[k] - stayed_k[k]), nl, 1);
            }
            if (KL == 16) {
#pragma unroll
                for (int k = 15; k >= 8; --k) {
                    ng2 = __funnelshift_l(__float_as_uint(stayed_k[k] - chg_k[k]), ng2, 1);
                    nl2 = __funnelshift_l(__float_as_uint(chg_k[k] - stayed_k[k]), nl2, 1);
                }
                *bp_ptr = (bp_t)((ng | (nl << 8)) | ((ng2 | (nl2 << 8)) << 16));
            } else {
                *bp_ptr = (bp_t)(ng | (nl << 8));
            }
            bp_ptr += bp_row;
            {   // first maximum of trellis[:, L] (:88); only the owner's comparison can be true
                unsigned vb = 0;
#pragma unroll
                for (int k = 0; k < KL; ++k) vb |= __float_as_uint(v[k]) & ownL[k];
                const float vl = __uint_as_float(vb);
                const bool up = ownsL && vl > best;
                best = up ? vl : best;
                best_t = up ? t0 + f + 1 : best_t;
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (ownsL) p.t_start[b] = best_t;
}

// Back-trace (:79-123) + merge_repeats (:141-157).  Warp 0 walks the "changed > stayed" bits from (t_start, L),
// 32 frames per round: lane i fetches the bits of frame t-1-i for the states [j-63, j] (issued one round ahead,
// the window covers wherever the current round ends) and compacts them into one 32-state window word; the walk
// itself is branch-free -- per frame one shuffle (off the dependent chain), a shift, a mask and a subtract -- and
// emits rec[frame] = (state << 1) | changed.  Then the whole CTA derives the token start frames, the per-frame
// probabilities (:106-112, which also need the "changed < stayed" bit of the path cell) and the per-token means.
// LK = 0: decision words of the barrier kernel (32 states per word, K "greater" then K "less" words per warp);
// LK = 4 / 8: lane entries of the wavefront kernel (16 states per 32-bit word, "greater" in the low half of
// every 2K-bit entry).
template <int LK>
__global__ void __launch_bounds__(1024) align_backtrace_kernel(const AlignParams p) {
    __shared__ int s_status, s_first;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    int Tb = p.em_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.Tmax ? (int)p.Tmax : Tb);
    int L = p.tok_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    const int NW = p.cfg.NW, KK = p.cfg.K;
    const int row_words = 2 * NW;  // 32-bit words of decision bits per frame (both layouts: 2 bits per state)
    const uint32_t *bp_b = p.bp + (int64_t)b * p.Tmax * row_words;
    uint32_t *rec = p.rec + (int64_t)b * p.Tmax;   // (state << 1) | (changed > stayed) for the frames on the path
    float *prob = p.prob + (int64_t)b * p.Tmax;
    int32_t *st_b = p.starts + (int64_t)b * p.Lmax;
    int32_t *en_b = p.ends + (int64_t)b * p.Lmax;
    double *sc_b = p.scores + (int64_t)b * p.Lmax;
    // status 2: the call's watchdog fired (see SpinGuard); status 3: a token id outside [0, V) -- the reference
    // would raise an IndexError (align_transcriptions.py:49), the kernels only clamp for memory safety
    __shared__ int s_bad;
    if (tid == 0) s_bad = (p.abort_word && *p.abort_word != kNoAbort) ? 2 : 0;
    __syncthreads();
    {
        const int32_t *tkc = p.tokens + (int64_t)b * p.tok_stride;
        int bad = 0;
        for (int i = tid; i < L; i += blockDim.x) bad |= (tkc[i] < 0 || tkc[i] >= p.V) ? 1 : 0;
        if (bad) atomicMax(&s_bad, 3);
    }
    __syncthreads();
    const int bad_status = s_bad;
    const int t_start = (L == 0 || Tb == 0 || bad_status) ? 0 : p.t_start[b];

    if (warp == 0) {
        int t = t_start, j = L, first = 0;
        // The decision bits of a long alignment come from HBM (C3: 1 GB), ~1 us away, and a round of 32 frames
        // walks in ~0.3 us: the bits are fetched DEPTH rounds ahead.  A fetch issued with the current state j covers
        // the states [j - 32 (DEPTH+1) + 1, j], wherever the walk will be DEPTH rounds later.
        constexpr int DEPTH = 3, SPAN = 32 * (DEPTH + 1);           // 128 states
        constexpr int NRAW = LK == 0 ? SPAN / 32 + 1 : SPAN / 16 + 1;  // 5 ballot words / 9 lane-entry words
        uint32_t cw[5], raw[DEPTH][NRAW];
        int cbase = 0, nbase[DEPTH] = {0, 0, 0};
        // issue the loads of the "changed > stayed" bits of trellis row tt-lane for the states from sb0 (a multiple
        // of 16 or 32) on; nothing here waits for them
        auto fetch = [&](int tt, int jj, uint32_t *w, int &sb0) {
            const int rr = tt - lane;
            const uint32_t *rowp = bp_b + (int64_t)(rr - 1) * row_words;
            if (LK == 0) {
                const int grp0 = max(jj - (SPAN - 1), 0) >> 5;
                sb0 = grp0 << 5;
#pragma unroll
                for (int q = 0; q < NRAW; ++q) {
                    const int grp = grp0 + q;
                    w[q] = (rr >= 1 && grp < NW) ? __ldg(rowp + (grp / KK) * 2 * KK + (grp % KK)) : 0u;
                }
            } else {
                const int w0 = max(jj - (SPAN - 1), 0) >> 4;
                sb0 = w0 << 4;
#pragma unroll
                for (int q = 0; q < NRAW; ++q) w[q] = (rr >= 1 && w0 + q < row_words) ? __ldg(rowp + w0 + q) : 0u;
            }
        };
        // raw words -> 160 consecutive state bits (first use of the loaded values)
        auto compact = [&](const uint32_t *w, uint32_t *g) {
            if (LK == 0) {
#pragma unroll
                for (int q = 0; q < 5; ++q) g[q] = w[q];
            } else {
                uint32_t c[10];
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    uint32_t x = w[q];
                    if (LK == 4) {
                        x &= 0x0f0f0f0fu;
                        x = (x | (x >> 4)) & 0x00ff00ffu;
                    } else {
                        x &= 0x00ff00ffu;
                    }
                    c[q] = (x | (x >> 8)) & 0xffffu;
                }
                c[9] = 0u;
#pragma unroll
                for (int q = 0; q < 5; ++q) g[q] = c[2 * q] | (c[2 * q + 1] << 16);
            }
        };
        // one round of 32 frames on the buffer `u` (compile-time index: the buffers rotate through an unrolled loop)
        auto round32 = [&](auto u_tag) {
            constexpr int U = decltype(u_tag)::value;
            compact(raw[U], cw);
            cbase = nbase[U];
            if (t > 32 * DEPTH) fetch(t - 32 * DEPTH, j, raw[U], nbase[U]);  // DEPTH rounds ahead, in flight meanwhile
            const int base = max(j - 31, 0);
            const int off = base - cbase;             // 0 <= off < 128
            const int kq = off >> 5;
            const uint32_t lo = kq == 0 ? cw[0] : (kq == 1 ? cw[1] : (kq == 2 ? cw[2] : cw[3]));
            const uint32_t hi = kq == 0 ? cw[1] : (kq == 1 ? cw[2] : (kq == 2 ? cw[3] : cw[4]));
            uint32_t win = __funnelshift_r(lo, hi, off & 31);
            if (base == 0) win &= ~1u;                // state 0 never "changes": the walk parks there (:119-120)
            const uint32_t valid = t >= 32 ? 0xffffffffu : ((1u << t) - 1u);  // step i looks at frame t-1-i >= 0
            // all 32 windows first (independent shuffles), then the dependent chain: shift, mask, subtract
            uint32_t gi[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t g = __shfl_sync(FULL, win, i);
                gi[i] = (valid >> i) & 1u ? g : 0u;
            }
            // dependent chain per frame: shift, mask, subtract; the decisions are collected in one word and every
            // lane reconstructs its own frame's state afterwards (state = start - number of earlier decisions)
            const int sh0 = j - base;                 // 0 <= sh <= 31, state = base + sh
            int sh = sh0;
            uint32_t dec = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t bit = (gi[i] >> sh) & 1u;  // :117 changed > stayed -> previous token
                dec += bit << i;
                sh -= (int)bit;
            }
            const int j_mine = base + sh0 - __popc(dec & ((1u << lane) - 1u));   // state at frame t-1-lane
            const uint32_t bit_mine = (dec >> lane) & 1u;
            const uint32_t mine = ((uint32_t)j_mine << 1) | bit_mine;
            // :119-120 the path starts at the frame whose decision leaves state 1 (only one lane can see that)
            const unsigned starts_here = __ballot_sync(FULL, bit_mine && j_mine == 1);
            if (starts_here) first = t - 1 - (__ffs(starts_here) - 1);
            j = base + sh;
            if (t - 1 - lane >= 0) rec[t - 1 - lane] = mine;
            t -= 32;
        };
        // prologue: the first DEPTH rounds are fetched with the start state (their windows still cover the walk)
        if (t > 0) fetch(t, j, raw[0], nbase[0]);
        if (t > 32) fetch(t - 32, j, raw[1], nbase[1]);
        if (t > 64) fetch(t - 64, j, raw[2], nbase[2]);
        while (t > 0 && j > 0) {
            round32(std::integral_constant<int, 0>{});
            if (!(t > 0 && j > 0)) break;
            round32(std::integral_constant<int, 1>{});
            if (!(t > 0 && j > 0)) break;
            round32(std::integral_constant<int, 2>{});
        }
        if (lane == 0) {
            const bool ok = j == 0 && L > 0 && t_start > 0;
            s_status = (ok && !bad_status) ? 0 : 1;  // :121-122 "Failed to align"
            s_first = first;
            p.status[b] = bad_status ? bad_status : (ok ? 0 : 1);
        }
    }
    __syncthreads();
    const bool ok = s_status == 0;
    const int first = ok ? s_first : 0, last = ok ? t_start : 0;   // path frames: [first, last)
    int32_t *ptok = p.path_token ? p.path_token + (int64_t)b * p.Tmax : nullptr;
    float *pprob = p.path_prob ? p.path_prob + (int64_t)b * p.Tmax : nullptr;
    const float *em_b = p.em + (int64_t)b * p.sb;
    const int32_t *tk = p.tokens + (int64_t)b * p.tok_stride;
    // per-frame probability of the path (:106-112) and the start frame of every token, one thread per frame
    for (int f = tid; f < (int)p.Tmax; f += blockDim.x) {
        float pr = 0.f;
        int ti = -1;
        if (f >= first && f < last) {
            const uint32_t rc = rec[f];
            const int j = (int)(rc >> 1);
            ti = j - 1;
            int tkn = tk[ti];
            tkn = tkn < 0 ? 0 : (tkn >= p.V ? p.V - 1 : tkn);
            const bool gt = rc & 1u;
            if (gt) st_b[ti] = f;
            // "changed < stayed" of the path cell
            const uint32_t *rowp = bp_b + (int64_t)f * row_words;
            bool lt;
            if (LK == 0) {
                const int grp = j >> 5;
                lt = (__ldg(rowp + (grp / KK) * 2 * KK + (grp % KK) + KK) >> (j & 31)) & 1u;
            } else if (LK == 4) {
                lt = (__ldg(rowp + (j >> 4)) >> (8 * ((j & 15) >> 2) + 4 + (j & 3))) & 1u;
            } else {
                lt = (__ldg(rowp + (j >> 4)) >> (16 * ((j & 15) >> 3) + 8 + (j & 7))) & 1u;
            }
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

struct AlignWs { size_t bp, rec, prob, col0, abort_word, gseam, gseam_bytes, total; };
static AlignWs align_ws_layout(int64_t B, int64_t Tmax, int nw, int S) {
    AlignWs w;
    const size_t bt = align_up((size_t)B * (size_t)Tmax * sizeof(float), 256);
    size_t o = 0;
    w.bp = o;    o += align_up((size_t)B * (size_t)Tmax * 2 * nw * sizeof(uint32_t), 256);
    w.rec = o;   o += bt;
    w.prob = o;  o += bt;
    w.col0 = o;  o += bt;
    w.abort_word = o;   // 256 bytes: the abort word, preset together with the seam buffer (one 0xff memset)
    o += 256;
    w.gseam = o;
    w.gseam_bytes = (size_t)B * (size_t)(S > 1 ? S - 1 : 0) * (size_t)Tmax * sizeof(float);
    o += align_up(w.gseam_bytes, 256);
    w.total = o + 256;
    return w;
}

extern "C" size_t ssak_align_workspace_bytes(int64_t B, int64_t Tmax, int64_t Lmax) {
    if (B <= 0 || Tmax < 0 || Lmax < 0) return 0;
    int S = 1;
    const int nw = align_max_nw(Lmax, B, &S);
    if (nw == 0) return 0;
    return align_ws_layout(B, Tmax, nw, S).total;
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
    // (the throughput kernel does not dump the trellis: a debugging request takes the other kernels)
    if (!choose_align_cfg(Lmax, B, (int)V, &p.cfg, trellis_dump == nullptr)) return SSAK_ERR_UNSUPPORTED;
    if (workspace_bytes < ssak_align_workspace_bytes(B, Tmax, Lmax)) return SSAK_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return SSAK_ERR_INVALID_ARGUMENT;
    const bool lane_mode = p.cfg.wave == 2;
    const bool wave = p.cfg.wave == 1;
    const size_t smem_bytes = lane_mode ? (size_t)lane_ring_bytes((int)V)
                              : wave ? (size_t)wave_smem(p.cfg.W, p.cfg.stages, p.cfg.chunk, p.cfg.slot_bytes).total
                                     : align_smem_bytes(p.cfg);
    if (smem_bytes > 227 * 1024) return SSAK_ERR_UNSUPPORTED;
    p.em = emissions; p.B = B; p.Tmax = Tmax; p.V = (int)V; p.sb = em_stride_b; p.st = em_stride_t;
    p.tokens = tokens; p.tok_stride = tok_stride; p.Lmax = (int)Lmax;
    p.em_len = emission_lengths; p.tok_len = token_lengths; p.blank = blank;
    p.garbage = first_as_garbage; p.col0 = col0;
    char *ws = reinterpret_cast<char *>(workspace);
    const AlignWs lay = align_ws_layout(B, Tmax, p.cfg.NW, p.cfg.S);
    p.bp = reinterpret_cast<uint32_t *>(ws + lay.bp);
    p.rec = reinterpret_cast<uint32_t *>(ws + lay.rec);
    p.prob = reinterpret_cast<float *>(ws + lay.prob);
    p.col0_eff = reinterpret_cast<float *>(ws + lay.col0);
    p.gseam = reinterpret_cast<float *>(ws + lay.gseam);
    p.abort_word = reinterpret_cast<int *>(ws + lay.abort_word);
    p.starts = starts; p.ends = ends; p.scores = scores; p.t_start = t_start; p.status = status;
    p.dump = trellis_dump; p.path_token = path_token; p.path_prob = path_prob;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int rc = SSAK_OK;
    {   // the abort word (and, for S > 1, every cross-CTA seam word: "not there yet") starts as all ones
        cudaError_t e = cudaMemsetAsync(p.abort_word, 0xff, 256 + (wave ? lay.gseam_bytes : 0), s);
        if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }
    }
    bool wave_launched = false;
    if (lane_mode) {
        if (p.cfg.W == 2) {
            if (V <= 64) align_lane_kernel<16, 2><<<(unsigned)B, 32, smem_bytes, s>>>(p);
            else align_lane_kernel<16, 4><<<(unsigned)B, 32, smem_bytes, s>>>(p);
        } else {
            if (V <= 64) align_lane_kernel<8, 2><<<(unsigned)B, 32, smem_bytes, s>>>(p);
            else align_lane_kernel<8, 4><<<(unsigned)B, 32, smem_bytes, s>>>(p);
        }
        wave_launched = true;
    } else if (wave && B <= 65535) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)p.cfg.S, (unsigned)B);
        lc.blockDim = dim3((p.cfg.W + 2) * 32);   // recursion warps, emission producer, column-0 warp
        lc.dynamicSmemBytes = smem_bytes;
        lc.stream = s;
        cudaLaunchAttribute attr[1];
        lc.attrs = attr;
        // The CTAs of an utterance poll each other's seams, so they MUST be co-resident: S > 1 is launched either
        // as a COOPERATIVE grid (the runtime guarantees that every CTA of the grid is resident, whatever else runs
        // on the device; chosen when the whole grid fits, because the block scheduler then spreads the CTAs one
        // per SM -- 8-CTA clusters fit only 15x on a B200) or as thread-block CLUSTERS of S <= 8 CTAs
        // (co-scheduled by the hardware).  If neither is possible the per-frame-barrier kernel (one CTA per
        // utterance) runs instead.  S == 1 has no cross-CTA dependency: plain launch.
        const int sms = device_sm_count();
        const bool prefer_cluster = align_env_int("SSAK_ALIGN_CLUSTER", 0) != 0;
#define SSAK_WAVE3(KK, CC, DD)                                                                 \
    {                                                                                          \
        auto kern = align_wave_kernel<KK, CC, DD>;                                             \
        cudaError_t e = ensure_max_smem<align_wave_kernel<KK, CC, DD>>();                      \
        int per_sm = 0;                                                                        \
        if (e == cudaSuccess)                                                                  \
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (int)lc.blockDim.x, smem_bytes); \
        if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }                \
        const bool fits = (int64_t)per_sm * sms >= (int64_t)p.cfg.S * B;                       \
        if (p.cfg.S == 1) {                                                                    \
            lc.numAttrs = 0;                                                                   \
            e = cudaLaunchKernelEx(&lc, kern, p);                                              \
            wave_launched = e == cudaSuccess;                                                  \
        } else {                                                                               \
            e = cudaErrorCooperativeLaunchTooLarge;                                            \
            if (fits && !prefer_cluster) {                                                     \
                attr[0].id = cudaLaunchAttributeCooperative;                                   \
                attr[0].val.cooperative = 1;                                                   \
                lc.numAttrs = 1;                                                               \
                e = cudaLaunchKernelEx(&lc, kern, p);                                          \
                wave_launched = e == cudaSuccess;                                              \
                if (!wave_launched) (void)cudaGetLastError();                                  \
            }                                                                                  \
            if (!wave_launched && p.cfg.S <= 8) {                                              \
                attr[0].id = cudaLaunchAttributeClusterDimension;                              \
                attr[0].val.clusterDim.x = (unsigned)p.cfg.S;                                  \
                attr[0].val.clusterDim.y = 1;                                                  \
                attr[0].val.clusterDim.z = 1;                                                  \
                lc.numAttrs = 1;                                                               \
                e = cudaLaunchKernelEx(&lc, kern, p);                                          \
                wave_launched = e == cudaSuccess;                                              \
                if (!wave_launched) (void)cudaGetLastError();                                  \
            }                                                                                  \
        }                                                                                      \
        if (!wave_launched && p.cfg.S == 1) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }  \
    }
#define SSAK_WAVE2(KK, CC)                                                                     \
    if (p.dump) SSAK_WAVE3(KK, CC, true) else SSAK_WAVE3(KK, CC, false)
#define SSAK_WAVE(KK)                                                                          \
    case KK:                                                                                   \
        if (p.cfg.chunk == 8) { SSAK_WAVE2(KK, 8) } else { SSAK_WAVE2(KK, 4) }                 \
        break;
        switch (p.cfg.K) {
            SSAK_WAVE(1)
            SSAK_WAVE(2)
            SSAK_WAVE(4)
            SSAK_WAVE(8)
            default: return SSAK_ERR_UNSUPPORTED;
        }
#undef SSAK_WAVE
#undef SSAK_WAVE2
#undef SSAK_WAVE3
    }
    if (!wave_launched) {
        // per-frame-barrier kernel: the configured shape, or the fall-back when the wavefront grid could not be
        // made co-resident (the workspace is sized for the larger of the two shapes)
        if (wave) {
            if (!choose_barrier_cfg(Lmax, B, (int)V, &p.cfg)) return SSAK_ERR_UNSUPPORTED;
            const AlignWs lay2 = align_ws_layout(B, Tmax, p.cfg.NW, p.cfg.S);
            p.bp = reinterpret_cast<uint32_t *>(ws + lay2.bp);
            p.rec = reinterpret_cast<uint32_t *>(ws + lay2.rec);
            p.prob = reinterpret_cast<float *>(ws + lay2.prob);
            p.col0_eff = reinterpret_cast<float *>(ws + lay2.col0);
            p.abort_word = reinterpret_cast<int *>(ws + lay2.abort_word);
            cudaError_t e = cudaMemsetAsync(p.abort_word, 0xff, 256, s);
            if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }
        }
        const size_t smem2 = align_smem_bytes(p.cfg);
        if (smem2 > (size_t)kMaxDynSmem) return SSAK_ERR_UNSUPPORTED;
        align_col0_kernel<<<(unsigned)B, 32, 0, s>>>(p);
        rc = check_launch();
        if (rc != SSAK_OK) return rc;
    dim3 grid((unsigned)B), block((p.cfg.W + 1) * 32);
#define SSAK_LAUNCH2(KK, CC)                                                                   \
    {                                                                                          \
        auto kern = align_forward_kernel<KK, CC>;                                              \
        cudaError_t e = ensure_max_smem<align_forward_kernel<KK, CC>>();                       \
        if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }                \
        kern<<<grid, block, smem2, s>>>(p);                                                    \
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
    }
    rc = check_launch();
    if (rc != SSAK_OK) return rc;
    // One warp walks; the whole CTA then turns the path into per-frame probabilities and per-token means -- a pass
    // of scattered loads over T frames.  Few long utterances (C3: 16 CTAs, 30000 frames each, 300 of the kernel's
    // 830 us in that pass): 1024 threads per CTA keep four times as many loads in flight.
    const unsigned bt_threads = (B <= device_sm_count() && Tmax >= 4096) ? 1024u : 256u;
    if (!wave_launched) align_backtrace_kernel<0><<<(unsigned)B, bt_threads, 0, s>>>(p);
    else if (p.cfg.K <= 4) align_backtrace_kernel<4><<<(unsigned)B, bt_threads, 0, s>>>(p);   // K = 1, 2 write the K = 4 layout
    else align_backtrace_kernel<8><<<(unsigned)B, bt_threads, 0, s>>>(p);
    return check_launch();
}
