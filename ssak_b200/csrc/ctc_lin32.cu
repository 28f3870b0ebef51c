// ctc_lin32.cu -- CTC loss forward / backward for sm_100a, throughput kernels: ONE WARP per (utterance, direction),
// linear domain, block floating point, no stored lattice.
//
// Same problem and the same meet-in-the-middle split as the log-domain kernels of ctc_loss.cu (replaces
// torch.nn.functional.ctc_loss as reached from ssak/train/transformers/wav2vec_train.py:313-325,
// ssak/train/speechbrain/wav2vec_train.py:66; arithmetic: SURVEY.md section 8 a-6 / a-7): chain (b,0) runs alpha over
// the frames [0,m), chain (b,1) runs beta over [m,T_b), a join kernel forms log P; the backward call continues both
// recursions over the other half, fused with the gradient.  What differs:
//   * LINEAR domain: per (blank,label) position and frame  A = b + carry;  t = l + b + skip*carry;  b' = A*y_blank;
//     l' = t*y_label -- with the blank state carried before its emission 3 FFMA + 1 FMUL and one shared-memory gather,
//     NO transcendental.  The emissions y = exp(lp) are formed once per (frame, vocabulary column) by the warp itself
//     (cp.async brings the raw rows into a shared-memory ring, one MUFU.EX2 per column converts them in place).
//   * ONE WARP owns the whole lattice row of its chain: lane j holds the K <= 13 consecutive positions
//     [jK, jK+K) in registers, a frame needs one shuffle (the lane-to-lane carry) and nothing else -- no barrier, no
//     cross-warp exchange, no mbarrier, no polling.  Warps are independent workers, ONE PER CTA (the block scheduler
//     hands an SM a new chain the moment one finishes; 12-16 are resident per SM), so the SM's four schedulers always
//     have ready warps: this is the throughput regime (hundreds of utterances), the opposite of the latency-tuned
//     kernels of ctc_loss.cu.
//   * BLOCK FLOATING POINT: fp32 mantissas with one integer exponent PER LANE.  Every C = 4 frames each lane moves
//     its largest state to ~2^TOP by an exact power of two; the lane exponents go through a decaying prefix-max
//     scan along the direction of flow (a lane is at most DMAX bits below its upstream neighbour, so the inflow
//     cannot overflow it), and the carry crossing a lane boundary is multiplied by 2^(E_up - E_me).  States more than
//     ~2^210 below their LANE's maximum flush to zero (-ftz): harmless unless such a state carries posterior mass,
//     which needs the forward and the backward partial likelihoods of one lane to disagree by that factor (garbage
//     transcripts).  Re-scaling has hysteresis (lane maxima stay in [2^84, 2^100)).  Rounding is relative (no
//     cancellation): the gradient is 2e-7 .. 4e-7 from the fp64 truth at T = 1500, ~300x closer than an fp32 log-domain
//     recursion (tools/proto_bfp.py is the numpy prototype the constants were chosen with).
//   * NOTHING of the lattice is written per frame.  forward() stores one CHECKPOINT row every C frames (1 B per
//     lattice cell; lanes far below the row maximum are dropped); backward() recomputes the C rows of a tile from its
//     checkpoint into a shared-memory tile private to the warp (already in posterior units: the recomputation starts
//     from the checkpoint times 2^(E_live + E_other - E_P + 30) / mantissa(P)), runs the live direction over the tile
//     and multiplies: posterior = live state before its emission x tile entry.  Label posteriors are added to a
//     per-warp mass vector in 2^-30 fixed point (shared-memory integer atomics: the sum does not depend on their
//     order), the blank ones through an integer warp reduction; the warp then writes the frame's gradient row,
//         grad[t,b,v] = (y[v] - mass[v] / sum_v mass[v]) * grad_out[b],
//     normalised by the frame's own total, so the common-mode drift of alpha_t beta_t against P cancels.
//   * SELF-CHECK and hand-back.  sum_s posterior_t(s) = 1 at every frame; states lost to the fp32 range can only
//     lower it.  A frame whose total is off by more than mass_tol (2e-5; rounding gives ~2e-6), a non-finite state,
//     a likelihood that is 0 for a feasible utterance, or emissions above 1 flag the utterance, and the log-domain
//     kernels of ctc_loss.cu recompute it (masked launches in the same call): the result never depends on the fp32
//     range.  (log P itself is carried to ~1e-6 absolute, like any fp32 recursion over 1500 frames.)
// Compiled with -ftz=true (denormals would only blur the flush threshold the self-check relies on).
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "ctc_lin32.h"

namespace ssak {
namespace lin32 {

constexpr int TOP = 96;             // re-scaled lane maximum ~ 2^TOP
constexpr int DEAD_BIT = 1 << 30;   // marker in a checkpoint's exponent word: the lane's states were dropped (all zero)
constexpr int ZDROP = 500;          // checkpoints drop lanes this many bits below the row's largest state
constexpr int WIN = 12;             // hysteresis of the re-scaling: lane maxima stay within [2^(TOP-WIN), 2^(TOP+4))
constexpr int DMAX = 16;            // a lane's exponent is at most DMAX below its upstream neighbour's
constexpr int EZERO = -(1 << 24);   // exponent wish of a lane that holds only zeros
// Copies of the posterior-mass vector (the lanes of a warp add into copy lane / (32 / copies)).  INTERLEAVED layout,
// mass[label][copy]: copy c only ever touches the banks = c (mod copies), so lanes of different copies never meet
// in a bank, the 32 / copies lanes of one copy spread over 32 / copies banks, and the frame's reduction reads the
// copies of a column as one or two 16-byte words.  8 copies for V <= 64, 4 beyond (shared memory per warp).
#ifndef SSAK_MASS_IL
#define SSAK_MASS_IL 1              // 0: the layout before (4 copies, mass[copy][label]) -- A/B builds
#endif
#ifndef SSAK_MASS_MC
#define SSAK_MASS_MC 8
#endif
__host__ __device__ inline int mc_of(int V) { return SSAK_MASS_IL ? (V <= 64 ? SSAK_MASS_MC : 4) : 4; }
constexpr int FIX = 30;             // posteriors are accumulated in 2^-FIX fixed point (integer adds: order-independent)
constexpr int GMIN = -120;          // smallest exponent of the tile scale 2^(E_live + E_other - E_P)
constexpr int GMAX = 127 - (TOP + DMAX + 2 * C + 2);   // largest exponent of tile scale x 2^FIX: tile entries stay finite
constexpr int FWD_WARPS = 16, BWD_WARPS = 12;   // resident one-warp CTAs per SM the register budgets are set for
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float pow2f(int e) {   // 2^e: 0 below 2^-126, 2^127 above
    e = e > 127 ? 127 : e;
    return e < -126 ? 0.f : __int_as_float((e + 127) << 23);
}

struct WarpSmem {
    int ring, tile, mass, total;   // byte offsets inside a warp's slice
};
constexpr int FWD_DEPTH = 4, BWD_DEPTH = 2;   // chunks / tiles of emission rows in shared memory (the one in use + look-ahead)
__host__ __device__ inline int nv_of(int V) { return V <= 64 ? 2 : 4; }      // vocabulary columns per lane (instantiated: 2, 4)
__host__ __device__ inline WarpSmem smem_map(int K, int V, bool grad) {
    WarpSmem m;
    int o = 0;
    m.ring = o;  o += (grad ? BWD_DEPTH : FWD_DEPTH) * C * (32 * nv_of(V) + 4) * 4;
    m.tile = o;  if (grad) o += C * 2 * K * 32 * 4;
    m.mass = o;  if (grad) o += (mc_of(V) * (V + 1) * 4 + 15) & ~15;
    m.total = (o + 127) & ~127;
    return m;
}

// The label state of position q = q0 + k is label q (D = 0) or label q - 1 (D = 1), and it exists iff that index is in
// [0, L): ONE table labs[i] = byte offset (4 x vocabulary index) of label q0 - 1 + i, i = 0 .. K, serves both
// directions (position k of direction D reads labs[k + 1 - D]); labels that do not exist point at the zero column
// of the emission ring / the dump slot of the posterior mass, offset 4 V.
template <int K>
__device__ __forceinline__ void make_labels(int q0, int L, int V, const int32_t *tg, int (&labs)[K + 1]) {
#pragma unroll
    for (int i = 0; i <= K; ++i) {
        const int li = q0 - 1 + i;
        labs[i] = 4 * V;
        if (li >= 0 && li < L) {
            const int l = tg[li];
            labs[i] = 4 * (l < 0 ? 0 : (l >= V ? V - 1 : l));   // (memory safety; the join kernel turns the likelihood into NaN)
        }
    }
}
// skip factor of position q: labels q-1 and q both exist and differ -- the same condition for alpha (label q may be
// entered from label q-1) and for beta (label q-1 may continue into label q)
template <int K>
__device__ __forceinline__ void make_skip(int q0, int L, int V, const int32_t *tg, float (&sk)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int q = q0 + k;
        sk[k] = 0.f;
        if (q >= 1 && q < L) {
            int l1 = tg[q - 1], l2 = tg[q];
            l1 = l1 < 0 ? 0 : (l1 >= V ? V - 1 : l1);
            l2 = l2 < 0 ? 0 : (l2 >= V ? V - 1 : l2);
            if (l1 != l2) sk[k] = 1.f;
        }
    }
}

// One frame of the recursion for the K positions of a lane.  The blank state is carried BEFORE its emission:
// a = (blank + carry) of the last frame, ebp = that frame's blank emission, true blank state = a * ebp, so that
//   A = a*ebp + carry;  t = (a*ebp + l) + skip*carry;  a' = A;  l' = t * y_label      (3 FFMA + 1 FMUL)
// Same arithmetic both ways, only the order of the lane's positions differs (the carry of a position is the OLD label
// state of its upstream neighbour).
//   MODE 0: plain.  MODE 1 (recomputation): the aligned copy of the row BEFORE this step goes to the tile -- the blank
//   state (as a; the consumer multiplies the blank sum by ebp once) and the label state the other direction pairs it
//   with, which is exactly the carry.  MODE 2 (live direction): posteriors = (state before its emission) x tile
//   entry, label posteriors added to mass[label] in fixed point.
template <int K, int D, int MODE, int MSH = 0>
__device__ __forceinline__ void step(float (&a)[K], float (&l)[K], const float ebp, const unsigned char *yrow,
                                     const int (&labs)[K + 1], const float (&sk)[K], const float cin, float *tl,
                                     unsigned *mass, float &sbl) {
    float carry = cin;
#pragma unroll
    for (int kk = 0; kk < K; ++kk) {
        const int k = D ? K - 1 - kk : kk;
        const float el = *reinterpret_cast<const float *>(yrow + labs[k + 1 - D]);
        const float A = fmaf(a[k], ebp, carry);
        const float t = fmaf(sk[k], carry, fmaf(a[k], ebp, l[k]));
        if (MODE == 1) {
            tl[k * 32] = a[k];
            tl[(K + k) * 32] = carry;
        }
        if (MODE == 2) {
            sbl = fmaf(A, tl[k * 32], sbl);
            atomicAdd(reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(mass) + (labs[k + 1 - D] << MSH)),
                      __float2uint_rn(t * tl[(K + k) * 32]));
        }
        carry = l[k];
        a[k] = A;
        l[k] = t * el;
    }
}

// lane-to-lane carry of direction D: the boundary label state of the upstream lane in MY scale (f = 0 at the chain's
// first lane)
template <int K, int D>
__device__ __forceinline__ float carry_in(const float (&l)[K], const float f) {
    const float nb = D ? __shfl_down_sync(FULL, l[0], 1) : __shfl_up_sync(FULL, l[K - 1], 1);
    return nb * f;
}

// Re-scale: every lane moves its largest state to ~2^TOP (exact powers of two), subject to
//   E_j >= Emin_j (the live direction of backward(): keeps the tile scale representable) and
//   E_j >= E_upstream - DMAX (decaying prefix-max scan along the flow: the inflow cannot overflow the lane).
// Returns false when a state is inf / NaN.
template <int K, int D>
__device__ __forceinline__ bool rescale(float (&a)[K], float (&l)[K], int &E, float &f, const int lane, const int Emin,
                                        const bool force, unsigned *hmax = nullptr) {
    unsigned h = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) h = max(h, max(__float_as_uint(a[k]), __float_as_uint(l[k])));   // states are >= 0
    if (hmax) *hmax = h;
    const int e_own = (int)(h >> 23) - 127;
    // hysteresis: a lane whose maximum sits in [2^(TOP-WIN), 2^(TOP+4)) and whose exponent respects its lower bound
    // needs nothing; most chunks end here (peaky emissions move a lane by ~0.3 bits per frame)
    const bool fine = h == 0u || (e_own >= TOP - WIN && e_own < TOP + 4 && E >= Emin);
    if (!force && __all_sync(FULL, fine)) return true;
    int v = h ? E + e_own - TOP : EZERO;
    v = max(v, Emin);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = D ? __shfl_down_sync(FULL, v, d) : __shfl_up_sync(FULL, v, d);
        const bool has = D ? lane + d < 32 : lane >= d;
        v = has ? max(v, o - d * DMAX) : v;
    }
    const float fac = pow2f(E - v);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        a[k] *= fac;
        l[k] *= fac;
    }
    if (hmax) *hmax = __float_as_uint(__uint_as_float(h) * fac);   // (the lane maximum after the shift)
    E = v;
    const int vu = D ? __shfl_down_sync(FULL, v, 1) : __shfl_up_sync(FULL, v, 1);
    f = lane == (D ? 31 : 0) ? 0.f : pow2f(vu - v);
    return h < 0x7f800000u;
}

template <int K>
__device__ __forceinline__ void start_row(int D, int q0, int L, float (&a)[K], float (&l)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        a[k] = (q0 + k == (D ? L : 0)) ? 1.f : 0.f;
        l[k] = 0.f;
    }
}
// checkpoint / frontier rows hold the TRUE states (blank = a * ebp): [2K][32] floats (blank k at k, label k at K+k),
// then the 32 lane exponents.  Lanes beyond the utterance's last position hold zeros and touch no memory.
template <int K>
__device__ __forceinline__ void store_row(float *row, const float (&a)[K], const float (&l)[K], float ebp, int E, int lane,
                                          bool live) {
    if (!live) return;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        __stcs(row + k * 32 + lane, a[k] * ebp);
        __stcs(row + (K + k) * 32 + lane, l[k]);
    }
    __stcs(reinterpret_cast<int *>(row) + 2 * K * 32 + lane, E);
}
// Checkpoints only.  A lane whose largest state is more than 2^ZDROP below the row's largest cannot carry posterior
// mass unless the other direction favours it by that factor (the self-check of backward() watches for exactly
// that): it is ZEROED -- in the registers too, so that forward() and the recomputation of backward() continue from
// the same row -- and only its exponent word is written, with bit 30 flipped as the marker (|E| < 2^29).  With peaky emissions ~3/4 of the
// lanes go this way: the checkpoint stream, which bounded forward() (0.69 ms with, 0.39 ms without it at B = 1024),
// shrinks accordingly.
template <int K>
__device__ __forceinline__ void store_checkpoint(float *row, float (&a)[K], float (&l)[K], float ebp, int E, unsigned h,
                                                 int lane, bool live) {
    const int x = h ? E + (int)(h >> 23) - 127 : EZERO;          // exponent of my largest state
    const int xmax = __reduce_max_sync(FULL, x);
    const bool dead = x < xmax - ZDROP;
    if (dead) {
#pragma unroll
        for (int k = 0; k < K; ++k) a[k] = l[k] = 0.f;
    }
    if (!live) return;
    if (!dead) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            __stcs(row + k * 32 + lane, a[k] * ebp);
            __stcs(row + (K + k) * 32 + lane, l[k]);
        }
    }
    // (the exponent itself is still needed: the lane keeps receiving its neighbour's carry in that scale)
    __stcs(reinterpret_cast<int *>(row) + 2 * K * 32 + lane, dead ? E ^ DEAD_BIT : E);
}
template <int K>
__device__ __forceinline__ void load_row(const float *row, float (&a)[K], float (&l)[K], int &E, int lane, bool live) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        a[k] = live ? __ldcs(row + k * 32 + lane) : 0.f;
        l[k] = live ? __ldcs(row + (K + k) * 32 + lane) : 0.f;
    }
    E = live ? __ldcs(reinterpret_cast<const int *>(row) + 2 * K * 32 + lane) : EZERO;
}
// a checkpoint lane that was dropped (exponent word with bit 30 flipped) left stale values in its slots: discard them
// and restore the exponent.  Called where the row is first USED -- the loads of load_row are issued a tile earlier
// and must not be waited for there.
template <int K>
__device__ __forceinline__ void drop_dead(float (&a)[K], float (&l)[K], int &E) {
    if (((E >> 30) ^ (E >> 31)) & 1) {
        E ^= DEAD_BIT;
#pragma unroll
        for (int k = 0; k < K; ++k) a[k] = l[k] = 0.f;
    }
}

struct Chain {
    int b, dir, Tb, L, m;
};
__device__ __forceinline__ bool chain_of(const Params &p, Chain &c) {
    const int64_t id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (id >= 2 * p.B) return false;
    if (p.order) {   // longest utterances first (see lin32_order_kernel)
        c.dir = (int)(id & 1);
        c.b = p.order[id >> 1];
    } else {
        c.dir = id >= p.B ? 1 : 0;
        c.b = (int)(id - (int64_t)c.dir * p.B);
    }
    int Tb = p.in_len[c.b];
    c.Tb = Tb < 0 ? 0 : (Tb > (int)p.T ? (int)p.T : Tb);
    int L = p.tgt_len[c.b];
    c.L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    c.m = c.Tb >> 1;
    return true;
}

// Launch order.  The block scheduler hands out one-warp CTAs in blockIdx order and an SM holds 12 (backward) chains:
// B = 1024 is 2048 chains on 1776 places, and the 272 chains of the partial second wave start when the first ones
// finish.  With the utterances in order of decreasing T_b the second wave holds the SHORTEST chains (and they run on
// nearly empty SMs): longest-processing-time-first.  order[rank of b] = b, rank = number of utterances with a larger
// T_b (ties by index): B^2 comparisons, one warp per utterance (3 us at B = 1024), B <= 8192.
constexpr int ORDER_MAX_B = 8192;
__global__ void __launch_bounds__(256) lin32_order_kernel(const Params p) {   // one warp per utterance
    const int lane = threadIdx.x & 31, b = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int B = (int)p.B;
    if (b >= B) return;
    auto len_of = [&](int i) {
        const int t = p.in_len[i];
        return t < 0 ? 0 : (t > (int)p.T ? (int)p.T : t);
    };
    const int mine = len_of(b);
    int rank = 0;
    for (int i = lane; i < B; i += 32) {
        const int t = len_of(i);
        rank += (t > mine || (t == mine && i < b)) ? 1 : 0;
    }
    rank = __reduce_add_sync(FULL, rank);
    if (lane == 0) p.order[rank] = b;
}

// Emission staging: the warp converts its own rows.  Raw rows travel global -> shared with cp.async (LDGSTS: no
// registers, no scoreboard -- completion is counted per commit group) into a ring of DEPTH chunks, DEPTH - 1 chunks
// ahead of their use; right before a chunk runs, its rows are converted (one MUFU.EX2 per column) into the warp's
// ring of C emission rows.  NV = columns per lane (V <= 32 NV); raw row stride = 32 NV + 4 floats (the spare ones
// hold the zero column [V] the non-existent states read, and the row normaliser of the logits entry points [RS-1]).
// The conversion happens IN PLACE: a ring row first holds log-probabilities, then emissions.
__host__ __device__ inline int raw_stride(int nv) { return 32 * nv + 4; }
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NV, int DEPTH>
struct Stage {
    static constexpr int RS = 32 * NV + 4;
    float *ring;      // [DEPTH][C][RS]
    __device__ __forceinline__ void init(int lane, int V) {   // the zero column of every row
        for (int r = lane; r < DEPTH * C; r += 32) ring[r * RS + V] = 0.f;
    }
    __device__ __forceinline__ const float *row(int s, int i) const { return ring + (size_t)(s * C + i) * RS; }
    // slot i of ring stage s <-> frame t0 + i*dt, slots [i0, nr); always commits a group (possibly empty)
    __device__ __forceinline__ void issue(int s, const float *lp_b, int64_t st, const float *zl_b, int64_t zstep, int lane,
                                          int V, int t0, int dt, int i0, int nr) {
        const float *src = lp_b + (int64_t)t0 * st + lane;
        const int64_t step = (int64_t)dt * st;
        float *dst = ring + (size_t)s * C * RS + lane;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            if (i >= i0 && i < nr) {
#pragma unroll
                for (int j = 0; j < NV; ++j)
                    if (lane + 32 * j < V) cp_async4(dst + i * RS + 32 * j, src + 32 * j);
                if (zl_b != nullptr && lane == 0) cp_async4(dst + i * RS + RS - 1, zl_b + (int64_t)(t0 + i * dt) * zstep);
            }
            src += step;
        }
        cp_async_commit();
    }
    // rows i0 .. nr-1 of ring stage s: log-probabilities -> emissions; returns the largest scaled emission seen
    // (log2 units; > 0: not a probability).  The caller has waited for the stage's group and synchronised the warp.
    __device__ __forceinline__ float convert(int s, int lane, int V, int i0, int nr, bool logits) {
        float *src = ring + (size_t)s * C * RS + lane;
        float xmax = -1.f;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            if (i >= i0 && i < nr) {
                const float z = logits ? src[i * RS + RS - 1 - lane] : 0.f;
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    if (lane + 32 * j < V) {
                        const float xs = fmaf(src[i * RS + 32 * j], kLog2e, z);
                        xmax = fmaxf(xmax, xs);
                        src[i * RS + 32 * j] = ex2_approx(xs);
                    }
                }
            }
        }
        return xmax;
    }
};

// ------------------------------------------------------------------------------------------------ forward
template <int K, int NV>
__global__ void __launch_bounds__(32, FWD_WARPS) lin32_forward_kernel(const Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Chain ch;
    if (!chain_of(p, ch)) return;
    const int b = ch.b, dir = ch.dir, Tb = ch.Tb, L = ch.L;
    const int V = p.V;
    const int nrows = dir ? Tb - ch.m : ch.m;
    const float *lp_b = p.lp + (int64_t)b * p.sb;
    const float *zl_b = p.zl ? p.zl + b : nullptr;
    const int32_t *tg = p.targets + p.tgt_off[b];
    const WarpSmem sm = smem_map(K, V, false);
    const int q0 = lane * K;
    const bool live = q0 <= L;

    auto run = [&](auto dtag) {
        constexpr int D = decltype(dtag)::value;
        int labs[K + 1];
        float sk[K], as[K], ls[K];
        make_labels<K>(q0, L, V, tg, labs);
        make_skip<K>(q0, L, V, tg, sk);
        start_row<K>(D, q0, L, as, ls);
        const int eoff_blank = 4 * p.blank;
        int E = 0;
        float f = 0.f, xmax = -1.f, dummy = 0.f, ebp = 1.f;
        bool ok = rescale<K, D>(as, ls, E, f, lane, EZERO, true);
        float *ck_dir = p.ck + ((int64_t)b * 2 + D) * p.NCK * p.ck_row;
        const int n_seq = (nrows + C - 1) / C;
        const int tf = D ? Tb - 1 : 0, dt = D ? -1 : 1;      // frame of row 0, direction of time
        Stage<NV, FWD_DEPTH> stg;
        stg.ring = reinterpret_cast<float *>(smem + (size_t)warp * sm.total + sm.ring);
        stg.init(lane, V);
        const bool logits = zl_b != nullptr;
        auto issue = [&](int n) {                            // chunk n into ring stage n % DEPTH (empty group beyond the end)
            const int left = nrows - n * C;
            stg.issue(n % FWD_DEPTH, lp_b, p.st, zl_b, p.B, lane, V, tf + dt * n * C, dt, 0, left < 0 ? 0 : (left < C ? left : C));
        };
#pragma unroll
        for (int n = 0; n < FWD_DEPTH - 1; ++n) issue(n);
        for (int n = 0; n < n_seq; ++n) {
            const int row0 = n * C, nr = nrows - row0 < C ? nrows - row0 : C;
            __syncwarp();                                    // every lane is done with the previous chunk's rows
            issue(n + FWD_DEPTH - 1);                        // (into the stage chunk n - 1 was converted from)
            cp_async_wait<FWD_DEPTH - 1>();                  // my copies of chunk n have landed ...
            __syncwarp();                                    // ... and so have everybody's
            xmax = fmaxf(xmax, stg.convert(n % FWD_DEPTH, lane, V, 0, nr, logits));
            __syncwarp();
            // (not unrolled over the frames: copies of the frame body per frame and direction overflow the
            //  instruction cache -- every warp is at its own place in the loop)
#pragma unroll 1
            for (int i = 0; i < nr; ++i) {
                const unsigned char *yrow = reinterpret_cast<const unsigned char *>(stg.row(n % FWD_DEPTH, i));
                const float eb = *reinterpret_cast<const float *>(yrow + eoff_blank);
                step<K, D, 0>(as, ls, ebp, yrow, labs, sk, carry_in<K, D>(ls, f), nullptr, nullptr, dummy);
                ebp = eb;
            }
            if (nr == C) {
                unsigned h;
                ok = rescale<K, D>(as, ls, E, f, lane, EZERO, false, &h) && ok;
                if (p.save) store_checkpoint<K>(ck_dir + (int64_t)(n + 1) * p.ck_row, as, ls, ebp, E, h, lane, live);
            }
        }
        cp_async_wait<0>();
        store_row<K>(p.fr + ((int64_t)b * 2 + D) * p.ck_row, as, ls, ebp, E, lane, live);
        // a partial last chunk is not re-scaled: check its states here
        unsigned h = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) h = max(h, max(__float_as_uint(as[k]), __float_as_uint(ls[k])));
        ok = ok && h < 0x7f800000u && !(xmax > 0.01f);
        if (__any_sync(FULL, !ok) && lane == 0) atomicOr(&p.flags[b], 1);   // inf / NaN / emissions above 1
    };
    if (dir) run(std::integral_constant<int, 1>{}); else run(std::integral_constant<int, 0>{});
}

// log P = log2( sum over the transitions from the alpha frontier (row m-1, or the virtual start row) into the beta
// frontier (row m) ) in fp64 (every state with its lane's exponent); feasibility (T_b >= L_b + repeats) decides
// between +inf and "let the log-domain kernels look at it".
__global__ void __launch_bounds__(128) lin32_join_kernel(const Params p) {
    __shared__ double red[4];
    __shared__ int redr[4], sEa[4], sEb[4];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int L = p.tgt_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    int Tb = p.in_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.T ? (int)p.T : Tb);
    const int K = p.K;
    const float *fa = p.fr + ((int64_t)b * 2) * p.ck_row, *fb = fa + p.ck_row;
    const int *ea = reinterpret_cast<const int *>(fa) + 64 * K, *eb = reinterpret_cast<const int *>(fb) + 64 * K;
    const int32_t *tg = p.targets + p.tgt_off[b];
    const int nl = L / K + 1;                                  // lanes that hold positions 0..L
    // reference exponents: the largest lane exponent of each frontier
    int Ea = EZERO * 2, Eb = EZERO * 2;
    for (int j = tid; j < nl; j += 128) {
        Ea = max(Ea, ea[j]);
        Eb = max(Eb, eb[j]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        Ea = max(Ea, __shfl_xor_sync(FULL, Ea, d));
        Eb = max(Eb, __shfl_xor_sync(FULL, Eb, d));
    }
    if (lane == 0) { sEa[warp] = Ea; sEb[warp] = Eb; }
    __syncthreads();
    Ea = max(max(sEa[0], sEa[1]), max(sEa[2], sEa[3]));
    Eb = max(max(sEb[0], sEb[1]), max(sEb[2], sEb[3]));
    auto val = [&](const float *row, const int *er, int Eref, bool label, int q) -> double {   // state of position q
        const int j = q / K, k = q - j * K;
        const int d = er[j] - Eref;
        return d < -1000 ? 0.0 : (double)row[((label ? K : 0) + k) * 32 + j] * exp2((double)d);
    };
    double sum = 0.0;
    int rep = 0, badt = 0;
    for (int q = tid; q <= L; q += 128) {
        const double ab = val(fa, ea, Ea, false, q);
        const double alp = q > 0 ? val(fa, ea, Ea, true, q - 1) : 0.0;
        sum += (ab + alp) * val(fb, eb, Eb, false, q);
        if (q < L) {
            const bool skip = q > 0 && tg[q] != tg[q - 1];
            const double t = val(fa, ea, Ea, true, q) + ab + (skip ? alp : 0.0);
            sum += t * val(fb, eb, Eb, true, q + 1);           // beta's label q sits at position q + 1
            rep += (q > 0 && tg[q] == tg[q - 1]) ? 1 : 0;
            badt |= (tg[q] < 0 || tg[q] >= p.V) ? 1 : 0;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_xor_sync(FULL, sum, d);
        rep += __shfl_xor_sync(FULL, rep, d);
    }
    if (lane == 0) { red[warp] = sum; redr[warp] = rep; }
    badt = __syncthreads_or(badt);
    if (tid == 0) {
        const double S = red[0] + red[1] + red[2] + red[3];
        const int R = redr[0] + redr[1] + redr[2] + redr[3];
        const bool feasible = Tb >= L + R;
        const float INF = __int_as_float(0x7f800000), QNAN = __int_as_float(0x7fc00000);
        if (badt) {
            p.nll[b] = QNAN;                                    // label outside the vocabulary (see ctc_join_kernel)
            p.nll2[b] = 0.0;
            p.flags[b] = 0;                                     // (decided here: nothing to hand back)
        } else if (!feasible) {
            p.nll[b] = INF;
            p.nll2[b] = 0.0;
            p.flags[b] = 0;
        } else {
            bool hand_back = (p.flags[b] & 1) != 0;             // set by the forward kernel (inf / NaN, emissions above 1)
            if (!(S > 0.0) || !(S < 1.0e300)) {
                hand_back = true;                               // zero, inf or NaN: the log-domain kernels decide
                p.nll[b] = QNAN;
                p.nll2[b] = 0.0;
            } else {
                const double logp2 = log2(S) + (double)Ea + (double)Eb;
                p.nll[b] = (float)(-logp2 * 0.6931471805599453);
                p.nll2[b] = -logp2;
            }
            if (hand_back) {
                // a row block for the log-domain kernels' stored half lattices; none left: NaN likelihood (loud)
                const int sl = atomicAdd(p.slot_counter, 1);
                if (sl < p.n_slots) {
                    p.slot[b] = sl;
                    p.slot_b[sl] = b;
                    p.flags[b] = 1;
                } else {
                    p.flags[b] = 4;                             // nobody owns it: backward() fills NaN
                    p.nll[b] = QNAN;
                }
            } else {
                p.flags[b] = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
// Chain (b, dir): the LIVE direction DL = dir continues from its frontier over the frames the other direction
// DR = 1 - dir covered in forward(); DR's rows are recomputed tile by tile from its checkpoints.  Tile j = rows
// [jC - 1, jC + C - 1) of DR (row -1 = the virtual start row, row jC - 1 = checkpoint j), processed j = J .. 0;
// ring slot / tile slot i <-> row jC - 1 + i.
template <int K, int NV>
__global__ void __launch_bounds__(32, BWD_WARPS) lin32_backward_kernel(const Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Chain ch;
    if (!chain_of(p, ch)) return;
    const int b = ch.b, dir = ch.dir, Tb = ch.Tb, L = ch.L;
    const int V = p.V;
    if (p.flags[b] & 5) return;                                // the log-domain kernels own this utterance (or nobody)
    const float nll = p.nll[b];
    const float gs = p.grad_out[b];
    {
        // trivial outcomes (as in ctc_lattice_kernel)
        const bool infeasible = nll == __int_as_float(0x7f800000);
        const bool isnan_ = nll != nll;
        if (infeasible || isnan_ || Tb == 0) {
            if (dir == 0) {
                const float fillv = (isnan_ || (infeasible && !p.zero_inf)) ? __int_as_float(0x7fc00000) : 0.f;
                for (int t = 0; t < (int)p.T; ++t) {
                    float *g = p.grad + (int64_t)t * p.gst + (int64_t)b * p.gsb;
                    const float v = t < Tb ? fillv : 0.f;
                    for (int cc = lane; cc < V; cc += 32) g[cc] = v;
                }
            }
            return;
        }
    }
    const int rdir = 1 - dir;
    const int nrows = rdir ? Tb - ch.m : ch.m;                 // rows of DR = frames this chain handles
    const float *lp_b = p.lp + (int64_t)b * p.sb;
    const float *zl_b = p.zl ? p.zl + b : nullptr;
    const int32_t *tg = p.targets + p.tgt_off[b];
    const WarpSmem sm = smem_map(K, V, true);
    unsigned char *mine = smem + (size_t)warp * sm.total;
    float *tile = reinterpret_cast<float *>(mine + sm.tile);       // [C][2K][32]
    // [V + 1][MC] label posterior mass in fixed point (integer adds: the sum does not depend on their order); MC
    // copies so that the ~L/V states of a label rarely meet in one shared-memory atomic (layout: see mc_of)
    unsigned *mass = reinterpret_cast<unsigned *>(mine + sm.mass);
    constexpr int MC = SSAK_MASS_IL ? (NV == 2 ? SSAK_MASS_MC : 4) : 4;      // (= mc_of(V): NV == 2 <=> V <= 64)
    constexpr int MSH = SSAK_MASS_IL ? (MC == 16 ? 4 : MC == 8 ? 3 : 2) : 0;   // label byte offset -> byte offset of its copies
    unsigned *mass_mine = SSAK_MASS_IL ? mass + lane / (32 / MC) : mass + (lane / (32 / MC)) * (V + 1);
    const int q0 = lane * K;
    const bool live = q0 <= L;
    for (int cc = lane; cc < MC * (V + 1); cc += 32) mass[cc] = 0u;

    auto run = [&](auto dtag) {
        constexpr int DL = decltype(dtag)::value, DR = 1 - DL;
        int labs[K + 1];
        float sk[K], la[K], ll[K], ra[K], rl[K];
        make_labels<K>(q0, L, V, tg, labs);
        make_skip<K>(q0, L, V, tg, sk);
        const int eoff_blank = 4 * p.blank;
        int EL = 0, ER = 0;
        float fL = 0.f, fR = 0.f, dummy = 0.f, ebpL = 1.f;
        load_row<K>(p.fr + ((int64_t)b * 2 + DL) * p.ck_row, la, ll, EL, lane, live);
        const double log2P = -p.nll2[b];
        const double Epd = floor(log2P);
        const float invPm = (float)exp2(Epd - log2P);           // 1 / mantissa of P, in (0.5, 1]
        const int Ep = (int)Epd;
        const float *ck_dir = p.ck + ((int64_t)b * 2 + DR) * p.NCK * p.ck_row;
        const int n_seq = nrows > 0 ? nrows / C + 1 : 0;
        const int dt = DR ? -1 : 1;
        auto geometry = [&](int n, int &j, int &row0, int &i0, int &nr) {
            j = nrows / C - n;
            row0 = j * C - 1;
            i0 = j == 0 ? 1 : 0;
            nr = nrows - row0 < C ? nrows - row0 : C;
        };
        auto frame_of = [&](int rho) { return DR ? Tb - 1 - rho : rho; };
        // checkpoint j of DR (j = 0: the virtual start row, re-scaled exactly as forward() did)
        auto load_ck = [&](int j) {
            if (j > 0) {
                load_row<K>(ck_dir + (int64_t)j * p.ck_row, ra, rl, ER, lane, live);
            } else {
                start_row<K>(DR, q0, L, ra, rl);
                ER = 0;
                float fdummy;
                rescale<K, DR>(ra, rl, ER, fdummy, lane, EZERO, true);
            }
        };
        Stage<NV, BWD_DEPTH> stg;
        stg.ring = reinterpret_cast<float *>(mine + sm.ring);
        stg.init(lane, V);
        const bool logits = zl_b != nullptr;
        bool bad = false;
        float xmax = -1.f;
        unsigned seen[NV];                                      // running sum of my columns' mass copies (see below)
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) seen[jj] = 0u;
        auto issue = [&](int n) {                               // tile n into ring stage n % DEPTH (empty group beyond the end)
            int j, row0, i0, nr;
            geometry(n, j, row0, i0, nr);
            if (n >= n_seq) nr = 0;
            stg.issue(n % BWD_DEPTH, lp_b, p.st, zl_b, p.B, lane, V, frame_of(row0), dt, i0, nr);
        };
#pragma unroll
        for (int n = 0; n < BWD_DEPTH - 1; ++n) issue(n);
        if (n_seq > 0) load_ck(nrows / C);
        for (int n = 0; n < n_seq; ++n) {
            int j, row0, i0, nr;
            geometry(n, j, row0, i0, nr);
            __syncwarp();                                       // the previous tile is done with the ring
            issue(n + BWD_DEPTH - 1);
            cp_async_wait<BWD_DEPTH - 1>();
            __syncwarp();
            xmax = fmaxf(xmax, stg.convert(n % BWD_DEPTH, lane, V, i0, nr, logits));
            // (L2 look-ahead for the checkpoint of the tile three tiles on: one instruction, no registers)
            if (j - 4 >= 1 && lane * 32 < p.ck_row)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ck_dir + (int64_t)(j - 4) * p.ck_row + lane * 32));
            __syncwarp();
            // ---- the live direction moves into this tile's scaling; the tile scale 2^(EL + ER - Ep) / mantissa(P)
            //      must be representable: EL >= GMIN - ER + Ep wherever DR holds anything
            drop_dead<K>(ra, rl, ER);                           // (first use of the checkpoint fetched a tile ago)
            {
                const int Emin = ER > EZERO / 2 ? GMIN - ER + Ep : EZERO;
                bad = !rescale<K, DL>(la, ll, EL, fL, lane, Emin, n == 0) || bad;
            }
            {
                // the recomputed rows go to the tile in posterior units (the recursion is linear: scale its start)
                int g = EL + ER - Ep + FIX;
                g = g > GMAX ? GMAX : (g < -127 ? -127 : g);
                const float gfac = pow2f(g) * invPm;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    ra[k] *= gfac;
                    rl[k] *= gfac;
                }
                // every lane now carries its own scale 2^(ER - g): the carry crossing a lane boundary converts between
                // the two (unclamped this is 2^(EL_me - EL_up): bounded by DMAX like the live direction's own carries)
                const int xr = ER - g;
                const int xu = DR ? __shfl_down_sync(FULL, xr, 1) : __shfl_up_sync(FULL, xr, 1);
                fR = lane == (DR ? 31 : 0) ? 0.f : pow2f(xu - xr);
            }
            // ---- R phase: rows row0 .. row0 + nr - 1 of DR into my column of the tile (aligned copies).  Slot 0
            //      holds true blank states (the checkpoint), the later slots a = blank / ebp: the B phase multiplies
            //      its blank sum by the frame's blank emission.  (Frame loops are not unrolled: see forward.)
            float *tl = tile + lane;                            // entry (i, kk) at (i * 2K + kk) * 32
            float ebpR = 1.f;
#pragma unroll 1
            for (int i = 0; i + 1 < nr; ++i) {
                const unsigned char *yrow = reinterpret_cast<const unsigned char *>(stg.row(n % BWD_DEPTH, i + 1));
                const float eb = *reinterpret_cast<const float *>(yrow + eoff_blank);
                step<K, DR, 1>(ra, rl, ebpR, yrow, labs, sk, carry_in<K, DR>(rl, fR), tl + i * 2 * K * 32, nullptr, dummy);
                ebpR = eb;
            }
            {
                // last row of the tile: only its aligned copy (the carries the next step would have fetched)
                const float cin = carry_in<K, DR>(rl, fR);
                float *ti = tl + (nr - 1) * 2 * K * 32;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    ti[k * 32] = ra[k];
                    const float c = DR ? (k == K - 1 ? cin : rl[k + 1 < K ? k + 1 : k]) : (k == 0 ? cin : rl[k > 0 ? k - 1 : 0]);
                    ti[(K + k) * 32] = c;
                }
            }
            // the next tile's checkpoint travels while the live direction works (ra / rl are dead until then; at 14
            // resident warps and 128 registers -- no room for this -- the kernel was 15 % slower)
            if (n + 1 < n_seq) load_ck(j - 1);
            // ---- B phase: the live direction over the tile's rows, last row first; every frame's posteriors go
            //      straight into its gradient row
#pragma unroll 1
            for (int i = nr - 1; i >= i0; --i) {
                const float *yrowf = stg.row(n % BWD_DEPTH, i);
                const unsigned char *yrow = reinterpret_cast<const unsigned char *>(yrowf);
                const float eb = *reinterpret_cast<const float *>(yrow + eoff_blank);
                float sbl = 0.f;
                step<K, DL, 2, MSH>(la, ll, ebpL, yrow, labs, sk, carry_in<K, DL>(ll, fL), tl + i * 2 * K * 32, mass_mine, sbl);
                ebpL = eb;
                sbl *= i == 0 ? 1.f : eb;                       // (slot 0: the checkpoint's true blank states)
                bad = bad || !(sbl <= 3.0e38f);                 // inf / NaN
                unsigned tot = __float2uint_rn(sbl);            // my share of the blank posterior, fixed point
                const unsigned blank_mass = __reduce_add_sync(FULL, tot);
                __syncwarp();                                   // the frame's label posteriors are in mass[]
                float *grow = p.grad + (int64_t)frame_of(row0 + i) * p.gst + (int64_t)b * p.gsb;
                unsigned mi[NV];
                tot = 0u;
#pragma unroll
                for (int jj = 0; jj < NV; ++jj) {
                    const int cc = lane + 32 * jj;
                    mi[jj] = 0u;
                    if (cc < V) {
                        unsigned msum = cc == p.blank ? blank_mass : 0u;
#if SSAK_MASS_IL
                        // the copies are never cleared: they run on (mod 2^32, like their sum) and the frame's mass
                        // is the difference to the sum the last frame saw
                        const uint4 *mp = reinterpret_cast<const uint4 *>(mass + cc * MC);
                        unsigned run = 0u;
#pragma unroll
                        for (int c2 = 0; c2 < MC / 4; ++c2) {
                            const uint4 v4 = mp[c2];
                            run += (v4.x + v4.y) + (v4.z + v4.w);
                        }
                        msum += run - seen[jj];
                        seen[jj] = run;
#else
#pragma unroll
                        for (int c2 = 0; c2 < MC; ++c2) {
                            msum += mass[c2 * (V + 1) + cc];
                            mass[c2 * (V + 1) + cc] = 0u;
                        }
#endif
                        mi[jj] = msum;
                    }
                    tot += mi[jj];
                }
                const float total = (float)__reduce_add_sync(FULL, tot);      // ~2^FIX
                bad = bad || !(fabsf(total * (1.0f / (float)(1 << FIX)) - 1.0f) <= p.mass_tol);
                const float inv = __fdividef(1.0f, total);
#pragma unroll
                for (int jj = 0; jj < NV; ++jj) {
                    const int cc = lane + 32 * jj;
                    if (cc < V) grow[cc] = (yrowf[cc] - (float)mi[jj] * inv) * gs;
                }
                __syncwarp();                                   // mass[] is clean for the next frame
            }
        }
        cp_async_wait<0>();
        bad = bad || xmax > 0.01f;
        if (__any_sync(FULL, bad) && lane == 0) {
            // hand the utterance to the log-domain kernels (they redo its forward in this call): first flagger
            // takes a row block
            const int old = atomicOr(&p.flags[b], 2);
            if (!(old & 2)) {
                const int sl = atomicAdd(p.slot_counter, 1);
                if (sl < p.n_slots) { p.slot[b] = sl; p.slot_b[sl] = b; } else atomicOr(&p.flags[b], 4);
            }
        }
    };
    if (dir) run(std::integral_constant<int, 1>{}); else run(std::integral_constant<int, 0>{});
    if (dir == 0) {   // frames beyond the utterance: exact zeros
        for (int t = Tb; t < (int)p.T; ++t) {
            float *g = p.grad + (int64_t)t * p.gst + (int64_t)b * p.gsb;
            for (int cc = lane; cc < V; cc += 32) g[cc] = 0.f;
        }
    }
}

// Utterances that were handed back when no row block was left (flag bit 2): nobody computed them -- NaN gradient
// (loud).  One warp per utterance; runs after everything else of backward().
__global__ void __launch_bounds__(256) lin32_orphan_kernel(const Params p) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= p.B || !(p.flags[b] & 4)) return;
    int Tb = p.in_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.T ? (int)p.T : Tb);
    for (int t = 0; t < (int)p.T; ++t) {
        float *g = p.grad + (int64_t)t * p.gst + b * p.gsb;
        const float v = t < Tb ? __int_as_float(0x7fc00000) : 0.f;
        for (int cc = lane; cc < p.V; cc += 32) g[cc] = v;
    }
}

}  // namespace lin32
}  // namespace ssak
#include "ctc_lin32_lv.cuh"
namespace ssak {
namespace lin32 {

// ------------------------------------------------------------------------------------------------ host side
bool ordered(int64_t B) { return B <= ORDER_MAX_B; }

int lanes_k(int64_t Lmax, int64_t V) {
    if (V > MAXV) {   // large vocabularies: the gather kernels of ctc_lin32_lv.cuh
        if (V > LV_MAXV || Lmax + 1 > 32 * LV_MAXK) return 0;
        return (Lmax + 1 + 31) / 32 <= 4 ? 4 : 7;
    }
    if (Lmax + 1 > 32 * MAXK) return 0;
    const int need = (int)((Lmax + 1 + 31) / 32);
    // at least C positions per lane: the inflow into a lane needs K frames to reach the next lane, so it cannot
    // cascade (and overflow) between two re-scalings
    return need <= 4 ? 4 : need <= 7 ? 7 : need <= 10 ? 10 : 13;
}

template <bool GRAD>
static int launch(const Params &p, cudaStream_t s) {
    // ONE WARP PER CTA: warps never cooperate, and the block scheduler then hands a new chain to an SM the moment one
    // finishes (12-warp CTAs ran 171 CTAs on 148 SMs in two waves).  Residency is bounded by registers and by the
    // warp's slice of shared memory (~15 KB in backward).
    const bool lv = p.V > MAXV;
    const size_t per_warp = lv ? (size_t)lv_smem_map(p.K, p.V, GRAD).total : (size_t)smem_map(p.K, p.V, GRAD).total;
    if (per_warp > (size_t)kMaxDynSmem) return SSAK_ERR_UNSUPPORTED;
    const int per_cta = 1;
    const unsigned grid = (unsigned)(2 * p.B);
    const size_t smem_bytes = per_cta * per_warp;
    if (lv) {
#define SSAK_LV(KK)                                                                                \
    {                                                                                              \
        cudaError_t e = GRAD ? ensure_max_smem<lv_backward_kernel<KK>>() : ensure_max_smem<lv_forward_kernel<KK>>(); \
        if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }                    \
        if (GRAD) lv_backward_kernel<KK><<<grid, 32, smem_bytes, s>>>(p);                          \
        else lv_forward_kernel<KK><<<grid, 32, smem_bytes, s>>>(p);                                \
    }
        if (p.K == 4) SSAK_LV(4) else if (p.K == 7) SSAK_LV(7) else return SSAK_ERR_UNSUPPORTED;
#undef SSAK_LV
        return check_launch();
    }
    const int nv = nv_of(p.V);
#define SSAK_L32B(KK, NN)                                                                          \
    {                                                                                              \
        if (GRAD) {                                                                                \
            cudaError_t e = ensure_max_smem<lin32_backward_kernel<KK, NN>>();                      \
            if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }                \
            lin32_backward_kernel<KK, NN><<<grid, per_cta * 32, smem_bytes, s>>>(p);               \
        } else {                                                                                   \
            cudaError_t e = ensure_max_smem<lin32_forward_kernel<KK, NN>>();                       \
            if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }                \
            lin32_forward_kernel<KK, NN><<<grid, per_cta * 32, smem_bytes, s>>>(p);                \
        }                                                                                          \
    }
#define SSAK_L32(KK) case KK: if (nv == 2) SSAK_L32B(KK, 2) else SSAK_L32B(KK, 4) break;
    switch (p.K) {
        SSAK_L32(4)
        SSAK_L32(7)
        SSAK_L32(10)
        SSAK_L32(13)
        default: return SSAK_ERR_UNSUPPORTED;
    }
#undef SSAK_L32
#undef SSAK_L32B
    return check_launch();
}

int launch_forward(const Params &p, cudaStream_t s) {
    if (p.order) {
        lin32_order_kernel<<<(unsigned)((p.B + 7) / 8), 256, 0, s>>>(p);
        int rc0 = check_launch();
        if (rc0 != SSAK_OK) return rc0;
    }
    int rc = launch<false>(p, s);
    if (rc != SSAK_OK) return rc;
    lin32_join_kernel<<<(unsigned)p.B, 128, 0, s>>>(p);
    return check_launch();
}

int launch_backward(const Params &p, cudaStream_t s) { return launch<true>(p, s); }

int launch_orphans(const Params &p, cudaStream_t s) {
    lin32_orphan_kernel<<<(unsigned)((p.B + 7) / 8), 256, 0, s>>>(p);
    return check_launch();
}

}  // namespace lin32
}  // namespace ssak
