// ctc_loss.cu -- CTC loss forward / backward for sm_100a (B200).
//
// Replaces torch.nn.functional.ctc_loss as the reference reaches it
// (ssak/train/transformers/wav2vec_train.py:313-325, ssak/train/speechbrain/wav2vec_train.py:66,
// ssak/train/nemo/yamls/model.yaml:3).  Arithmetic follows SURVEY.md section 8 a-6 / a-7.
//
// Design (B200-first, not a port of ATen's LossCTC.cu):
//   * "Meet in the middle".  For utterance b with T_b frames, one CTA runs the alpha
//     recursion over frames [0, m) and a second CTA runs the beta recursion over frames
//     [m, T_b), m = T_b/2, AT THE SAME TIME (grid = B x 2).  The beta recursion is the alpha
//     recursion of the time-reversed, label-reversed problem, so both CTAs execute the same
//     code.  A small join kernel combines the two frontier rows into the log-likelihood.
//     The backward call resumes both recursions over the other half of the frames and fuses
//     the gradient: the CTA continuing alpha over [m, T_b) reads the beta rows the forward
//     call stored, and vice versa.  Serial depth per call is T/2 instead of T, stored
//     lattice traffic is 4 B/cell written + 4 B/cell read (half of a full alpha + full beta).
//   * The 2L+1 extended states are handled as L+1 (blank, label) PAIRS.  With
//     A = lse(alpha[2p], alpha[2p-1]) the blank update is A + lp[blank] and the label update is
//     lse(alpha[2p+1], skip ? A : alpha[2p]) + lp[label]: 2 log-sum-exp of two terms per pair,
//     i.e. 2 MUFU per lattice cell instead of 3-4, all in the log2 domain (ex2/lg2 are the
//     native MUFU ops; emissions are scaled by log2(e) when gathered).
//   * Pairs are spread cyclically over the lanes of a warp (pair = warp*32K + k*32 + lane), so
//     the neighbour state comes from one lane rotation (warp shuffle) per k, every global row
//     access is a coalesced 128-byte line at an immediate offset, and only one value per warp
//     crosses warps per frame (shared memory, double buffered, one named barrier per frame
//     among the recursion/gradient warps).  Both directions keep states in natural order; the
//     beta CTA pairs (label p-1, blank p) and rotates lanes the other way.  States beyond the
//     utterance's 2L+1 are pinned to log(0) by a sentinel emission slot, so the time loop has
//     no validity predicates.
//   * A dedicated producer warp, decoupled from the per-frame barrier (mbarriers only),
//     prefetches the emission rows of the next frames into a shared-memory ring with
//     cp.async.bulk (1-D TMA); in the backward call it also streams the other direction's
//     stored lattice rows into a second ring.  Nothing but the copy engine touches the rows.
//   * Numerics: every 8 frames the row is re-centred on its maximum (one integer REDUX per
//     warp) and the subtracted amount is accumulated in fp64, so the fp32 state values stay
//     O(10..100) instead of O(T): the rounding noise of the recursion drops by ~100x
//     compared with a plain fp32 log-domain recursion (what ATen does).
//   * Backward: dedicated gradient warps run behind the recursion warps.  Per-state posteriors are published in
//     shared memory, blank posteriors are summed with an integer warp reduction in 2^-30 fixed point
//     (deterministic), label posteriors are summed by the gradient warps through a per-utterance label-sorted
//     order built once, and full gradient rows are written coalesced.  In the latency regime (few CTAs) the
//     posteriors are computed by POSTERIOR WARPS, one per recursion warp and one chunk behind it: the recursion
//     warp hands over its states before the emission is added (a 2-chunk state ring), so the serial chain carries
//     nothing but the recursion (a lone warp issues ~0.3 instructions per cycle whatever its ILP).
//   * Forward in the latency regime: ctc_forward_wave_kernel -- no per-frame barrier, warps skewed in time,
//     per-warp integer re-centring with side tables of offsets (see "offset tables" below).
//   * Logits entry points: ctc_row_lse_kernel + the LOGITS flag of the lattice kernels (log_softmax never
//     materialised).  Sharded (multi-GPU) reduction: ctc_shard_*_kernel around the caller's one all-reduce.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "ctc_lin32.h"

namespace ssak {


struct CtcCfg {
    int K;          // pairs per lane
    int W;          // recursion warps
    int G;          // gradient warps (backward kernel only)
    int P_pad;      // pair capacity = 32*K*W
    int row_elems;  // floats per stored lattice row (see "row layout")
    int chunk;      // frames per emission ring stage
    int stages;     // emission ring stages
    int slot_bytes;
    int or_chunk;   // rows per stage of the "other direction" ring (backward)
    int or_stages;
    int PW;         // posterior warps (backward; 0: the recursion warps compute the posteriors themselves)
};
// Row layout (floats): [0,P_pad) blank states | [P_pad] spare | [P_pad+1, 2P_pad+1) label states
//                      | [2P_pad+2, 2P_pad+4) fp64 re-centring offset | pad to 2P_pad+8.
// [2P_pad+4] of a frontier row holds Kf: 0 when the barrier forward wrote the rows, else the chain elements per
// lane of the wavefront forward.
// Offset tables (Kf > 0 only).  The wavefront forward re-centres per warp, by integers: the states of its warp w
// (chain elements [32 Kf w, 32 Kf (w+1)); element c = pair c for alpha, pair L-c for beta) are relative to
// row offset + table[w], table = tabs[b][dir][forward chunk of the frame] (frontier: slot NCH-1).  The stored
// states thus stay small wherever the probability mass sits; consumers add the integers back exactly.
// Label p lives at P_pad+1+p.  The beta CTA's thread for pair q owns (label q-1, blank q), i.e.
// positions (P_pad+q, q): both directions use immediate offsets from one per-thread base.

struct CtcParams {
    const float *lp;
    int64_t T, B;
    int V;
    int64_t st, sb;
    const int32_t *targets;
    const int64_t *tgt_off;
    const int32_t *in_len;
    const int32_t *tgt_len;
    int Lmax;
    int blank;
    // workspace
    double *nll2;   // [B]    -log2 P (fp64: sum of the offsets + joined frontier)
    float *finals;  // [B][2][row_elems] frontier rows
    float *rows;    // [B][T][row_elems] half lattices (saved for backward)
    float *zl;      // [T][B] -log2(sum_v exp(x[t,b,v])) when the input holds raw logits, else nullptr
    float *tabs;    // [B][2][NCH][16] per-warp offset tables of the wavefront forward (see "offset tables")
    int NCH;        // table slots per (utterance, direction): one per forward chunk, the last one for the frontier
    float *nll;     // [B] out / in
    int *abort_word;  // 0 until a seam poll of the wavefront forward gave up (watchdog); then every nll of the call is NaN
    const int *slot;  // [B] or nullptr: row block of utterance b in `rows` (throughput mode: only the utterances handed
                      // back get one, see ctc_lin32.cu); nullptr: block b
    const int *mask;  // [B] or nullptr: when set, only the utterances with one of `mask_bits` set (and bit 2 clear) are
                      // processed: the ones the throughput kernels of ctc_lin32.cu handed back (flags, see there)
    int mask_bits;
    const int *slot_b;      // [n_slots] utterance that owns row block i: a masked launch has one CTA pair per ROW BLOCK
    const int *slot_count;  // row blocks handed out so far (may exceed n_slots: the late ones got none)
    int n_slots;
    int join_keeps_nll;   // the join kernel leaves nll[b] alone (backward-time recomputation: nll is the caller's input)
    int *nan_flag;    // [B] set by the forward kernels when an emission the lattice uses is NaN (fmax-based log-sum-exp
                      // would swallow it): the join kernel then returns a NaN likelihood, as torch does
    const float *grad_out;
    float *grad;
    int64_t gst, gsb;
    int zero_inf;
    CtcCfg cfg;
};

static inline int env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Launch shape from (Lmax, B) only, so forward and backward agree on the workspace layout.
static size_t smem_bytes_for(const CtcCfg &c, int V, int Lmax, bool grad);
static bool choose_cfg(int64_t Lmax, int64_t B, int V, CtcCfg *c) {
    const int64_t P = Lmax + 1;
    // Few CTAs (latency regime): one recursion warp per SM sub-partition.  Many CTAs
    // (throughput regime): fat lanes, few warps, so several utterances share an SM.
    const int sms = device_sm_count();
    const bool few = env_int("SSAK_CTC_FEW", B <= sms ? 1 : 0) != 0;
    int wtarget = few ? 8 : 4;
    wtarget = env_int("SSAK_CTC_WARPS", wtarget);
    int K = env_int("SSAK_CTC_K", 0);
    if (K == 0) {
        K = 1;
        while (K < 4 && (P + 32 * K - 1) / (32 * K) > wtarget) K *= 2;
    }
    if (K != 1 && K != 2 && K != 4 && K != 8) return false;
    if (K < 8 && (P + 32 * K - 1) / (32 * K) > 8) K = 8;  // K <= 4 kernels are built for <= 8 recursion warps
    const int64_t W = (P + 32 * K - 1) / (32 * K);
    if (W > 16) return false;  // L <= 4095
    c->K = K;
    c->W = (int)W;
    c->P_pad = 32 * K * (int)W;
    c->row_elems = 2 * c->P_pad + 8;
    // gradient warps: few when many CTAs share an SM (they only cost occupancy), more when one CTA owns it
    c->G = few ? (V <= 512 ? 4 : 8) : (V <= 512 ? 2 : 4);
    if (K == 8 && c->W + 2 + c->G > 22) c->G = 22 - 2 - c->W;
    // (set below once the chunk is known: G <= chunk, every gradient warp owns a frame of every chunk)
    c->slot_bytes = ring_slot_bytes(V);
    c->chunk = 8 * c->slot_bytes <= 16384 ? 8 : 4;   // kernels are instantiated for 8 and 4
    if (c->G > c->chunk) c->G = c->chunk;            // every gradient warp owns a frame of every chunk
    // >= 3 stages: two chunks of look-ahead are needed to cover the HBM latency of the bulk copies
    int stages = (64 * 1024) / (c->chunk * c->slot_bytes);
    c->stages = stages > 4 ? 4 : stages;
    if (c->stages < 2) return false;                 // V too large for the emission ring
    const int row_bytes = c->row_elems * 4;
    // other-direction rows (backward): chunks of 2 rows, as many stages as the budget allows (>= 8 rows
    // of look-ahead when they fit: one row is consumed per frame and a bulk copy takes ~1-2 us)
    const int budget = few ? 96 * 1024 : 40 * 1024;
    // few CTAs (one per SM): big chunks (fewer barrier probes on the critical path); many CTAs: 2-row chunks
    int oc = few ? 8 : 2;
    while (oc > 1 && oc * 3 * row_bytes > budget) oc >>= 1;
    int ost = budget / (oc * row_bytes);
    c->or_chunk = oc;
    c->or_stages = ost < 2 ? 2 : (ost > (few ? 3 : 8) ? (few ? 3 : 8) : ost);
    c->G = env_int("SSAK_CTC_G", c->G);
    // (on request only: 2*chunk gradient warps, pairs sharing a frame and taking half of its columns each --
    //  measured slower at C5, 1.33 vs 0.85 ms: the extra warps cost resident CTAs)
    if (c->G > c->chunk) c->G = (c->G >= 2 * c->chunk && V >= 256) ? 2 * c->chunk : c->chunk;
    c->or_chunk = env_int("SSAK_CTC_OR_CHUNK", c->or_chunk);
    c->or_stages = env_int("SSAK_CTC_OR_STAGES", c->or_stages);
    if (c->or_stages > 8 || c->or_stages < 2 || c->or_chunk < 1) return false;
    // Posterior warps (latency regime only): a lone warp issues ~0.3 instructions per cycle whatever its ILP, so
    // the per-frame work of the backward is split over two warps per state group -- the recursion warp keeps the
    // serial chain, a posterior warp (one chunk behind, through a 2-chunk state ring) multiplies with the other
    // direction's row.  Needs W more warps (<= 22 in all) and 2*chunk*2*P_pad floats of shared memory.
    c->PW = 0;
    if (few && env_int("SSAK_CTC_SPLIT", 1) != 0 && K <= 4 && c->W + 2 + c->G + c->W <= 22) {
        c->PW = c->W;
        if (smem_bytes_for(*c, V, (int)Lmax, true) > 227 * 1024) c->PW = 0;
    }
    return true;
}

// shared memory map (bytes)
constexpr int kBarEmFull = 0, kBarEmEmpty = 64, kBarOrFull = 128, kBarOrEmpty = 192, kBarPostEmpty = 256, kBarPostFull = 272;
constexpr int kSmemXchg = 400, kSmemWmax = 560, kSmemRing = 640;
// bytes of the posterior-warp hand-over area (state ring of 2 chunk buffers, per-warp base offsets, mbarriers)
static size_t split_bytes_for(const CtcCfg &c) {
    return c.PW ? 2 * (size_t)c.chunk * 2 * c.P_pad * sizeof(float) + (size_t)c.PW * 2 * 8 + (size_t)c.PW * 4 * 8 : 0;
}
static size_t smem_bytes_for(const CtcCfg &c, int V, int Lmax, bool grad) {
    size_t o = kSmemRing + (size_t)c.stages * c.chunk * c.slot_bytes;
    if (grad) {
        o += (size_t)c.or_stages * c.or_chunk * c.row_elems * 4 + 2 * (size_t)c.chunk * (c.P_pad + 8) * sizeof(float) +
             2 * (size_t)c.chunk * sizeof(unsigned) + ((size_t)V + 2) * sizeof(int) + (size_t)V * sizeof(int) +
             (size_t)(Lmax > 0 ? Lmax : 1) * sizeof(int);
        o = align_up(o, 16) + split_bytes_for(c);
    }
    return align_up(o, 16);
}

// The utterance of this CTA: blockIdx.x -- or, in a masked launch (grid = row blocks, not utterances: a launch over
// all B x 2 CTAs cost ~100 us at B = 1024 just to find out that nothing was handed back), the utterance that owns
// row block blockIdx.x; -1: nothing to do.
__device__ __forceinline__ int utterance_of_cta(const CtcParams &p) {
    int b = blockIdx.x;
    if (p.mask) {
        if (b >= *p.slot_count) return -1;
        b = p.slot_b[b];
        const int fl = p.mask[b];
        if (!(fl & p.mask_bits) || (fl & 4)) return -1;
    }
    return b;
}

// ------------------------------------------------------------------------------ kernel
// Warp roles: [0, W) recursion; W emission producer; backward only: W+1 lattice-row producer, then G gradient warps.
template <int K, bool GRAD, int CH, bool LOGITS, bool SPLIT>
__global__ void __launch_bounds__(K == 8 ? (GRAD ? 704 : 544) : (GRAD ? (SPLIT ? 704 : 576) : 288), 1)
ctc_lattice_kernel(const CtcParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const CtcCfg &c = p.cfg;
    const int b = utterance_of_cta(p);
    if (b < 0) return;
    const int dir = blockIdx.y;  // 0: alpha (forward in time), 1: beta (backward in time)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = c.W;
    constexpr int NPROD = GRAD ? 2 : 1;  // producer warps
    const bool producer = warp >= W && warp < W + NPROD;
    const int PW = (GRAD && SPLIT) ? c.PW : 0;           // posterior warps (0: the recursion warps do it)
    const int post0 = W + NPROD + (GRAD ? c.G : 0);      // first posterior warp
    const bool is_post = SPLIT && warp >= post0;
    constexpr bool split = GRAD && SPLIT;
    const int rw = is_post ? warp - post0 : warp;        // the recursion warp whose state group this warp handles
    const unsigned FULL = 0xffffffffu;

    int Tb = p.in_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.T ? (int)p.T : Tb);
    int L = p.tgt_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    const int m = Tb >> 1;
    // recursion warps (and their posterior warps) whose 32K pairs all lie beyond pair L of THIS utterance have
    // nothing to compute: they leave before the time loop, the barriers below count the live ones only
    const int wl = (L / (32 * K) + 1) < W ? (L / (32 * K) + 1) : W;
    const int n1 = dir ? Tb - m : m;        // frames this direction owns in forward()
    const int tau0 = GRAD ? n1 : 0;         // first direction-local step of this launch
    const int nsteps = GRAD ? Tb - n1 : n1;
    const int P_pad = c.P_pad;
    const int V = p.V;
    const int row_elems = c.row_elems;
    const float *lp_b = p.lp + (int64_t)b * p.sb;
    const int32_t *tg = p.targets + p.tgt_off[b];
    const int t_first = dir ? Tb - 1 - tau0 : tau0;  // frame of step 0
    const int dt = dir ? -1 : 1;
    const bool compute = warp < wl;
    const bool dead = (warp < W && warp >= wl) || (is_post && rw >= wl);

    uint64_t *em_full = reinterpret_cast<uint64_t *>(smem + kBarEmFull);
    uint64_t *em_empty = reinterpret_cast<uint64_t *>(smem + kBarEmEmpty);
    uint64_t *or_full = reinterpret_cast<uint64_t *>(smem + kBarOrFull);
    uint64_t *or_empty = reinterpret_cast<uint64_t *>(smem + kBarOrEmpty);
    uint64_t *post_empty = reinterpret_cast<uint64_t *>(smem + kBarPostEmpty);  // [2]  per chunk buffer
    uint64_t *post_full = reinterpret_cast<uint64_t *>(smem + kBarPostFull);    // [2*CH] per frame slot
    float *xchg = reinterpret_cast<float *>(smem + kSmemXchg);              // [2][18]: guard, W seams, guard
    float *wmax = reinterpret_cast<float *>(smem + kSmemWmax);              // [16]
    RowRing ring;
    ring.slots = smem + kSmemRing;
    ring.full = em_full;
    ring.chunk = c.chunk;
    ring.stages = c.stages;
    ring.slot_bytes = c.slot_bytes;
    ring.row_bytes = 4 * V;
    unsigned char *or_slots = smem + kSmemRing + (size_t)c.stages * c.chunk * c.slot_bytes;
    const int row_bytes = row_elems * 4;
    const int WL = P_pad + 8;
    // posterior ring (backward): 2 chunk buffers x CH frames, label posteriors in label-sorted order
    float *wlab = reinterpret_cast<float *>(or_slots + (size_t)c.or_stages * c.or_chunk * row_bytes);  // [2*CH][WL]
    unsigned *blank_acc = reinterpret_cast<unsigned *>(wlab + 2 * CH * WL);                            // [2*CH]
    int *occ_start = reinterpret_cast<int *>(blank_acc + 2 * CH);
    int *cursor = occ_start + (V + 2);  // counting-sort cursors, then the list of columns that carry posterior mass
    int *occ_pos = cursor + V;
    // posterior-warp hand-over (split): state ring [2*CH][2*P_pad], per-warp chunk bases, mbarriers
    unsigned char *split_base = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(occ_pos + (p.Lmax > 0 ? p.Lmax : 1)) + 15) & ~(uintptr_t)15);
    float *sring = reinterpret_cast<float *>(split_base);
    double *sbase = reinterpret_cast<double *>(split_base + 2 * (size_t)CH * 2 * P_pad * sizeof(float));  // [PW][2]
    uint64_t *st_full = reinterpret_cast<uint64_t *>(sbase + 2 * (PW > 0 ? PW : 1));                       // [PW][2]
    uint64_t *st_empty = st_full + 2 * (PW > 0 ? PW : 1);                                                  // [PW][2]

    // ---- gradient prologue: trivial outcomes ----
    double nll2 = 0.0;
    float gs = 0.f;
    if (GRAD) {
        const float nll = p.nll[b];
        gs = p.grad_out[b];
        const bool infeasible = nll == __int_as_float(0x7f800000);  // +inf only: a NaN stays visible, as in torch
        const bool isnan_ = nll != nll;
        if (infeasible || isnan_ || Tb == 0) {
            // zero_infinity: every row of an infeasible utterance 0.  Otherwise torch yields NaN for t < T_b; a NaN
            // likelihood (NaN emissions of a diverged model, or an out-of-range target) gives a NaN gradient always.
            if (dir == 0) {
                const float fillv = (isnan_ || (infeasible && !p.zero_inf)) ? __int_as_float(0x7fc00000) : 0.f;
                for (int t = 0; t < (int)p.T; ++t) {
                    float *g = p.grad + (int64_t)t * p.gst + (int64_t)b * p.gsb;
                    const float v = t < Tb ? fillv : 0.f;
                    for (int cc = tid; cc < V; cc += blockDim.x) g[cc] = v;
                }
            }
            return;
        }
        nll2 = p.nll2[b];
    }

    const int n_consumers = W + (GRAD ? c.G : 0) + PW;   // every warp but the producers
    const int n_em = wl + (GRAD ? c.G : 0);              // warps that read the emission ring
    if (tid == 0) {
        for (int s = 0; s < c.stages; ++s) {
            mbar_init(&em_full[s], 1);
            mbar_init(&em_empty[s], n_em);
        }
        for (int s = 0; s < c.or_stages; ++s) {
            mbar_init(&or_full[s], 1);
            mbar_init(&or_empty[s], wl);
        }
        if (GRAD) {
            for (int s = 0; s < 2 * CH; ++s) {
                mbar_init(&post_full[s], wl);
                blank_acc[s] = 0u;
            }
            mbar_init(&post_empty[0], c.G);
            mbar_init(&post_empty[1], c.G);
            for (int s = 0; s < 2 * PW; ++s) {
                mbar_init(&st_full[s], 1);
                mbar_init(&st_empty[s], 1);
            }
        }
        mbar_fence_init();
    }
    if (tid < 36) xchg[tid] = kNeg;  // seam guards (and every seam until its warp writes it)
    // Sentinel emission log(0) for the states beyond 2L+1: the last 16 bytes of every ring slot are never
    // written by the bulk copies (a row lands within the first slot_bytes-16 bytes).
    for (int i = tid; i < c.stages * CH * 4; i += blockDim.x)
        *reinterpret_cast<float *>(ring.slots + (size_t)(i >> 2) * c.slot_bytes + c.slot_bytes - 16 + (i & 3) * 4) = kNeg;
    __syncthreads();

    // ---- per-thread static data: emission byte offsets of my K labels, skip flags ----
    // alpha: pair p = (blank p, label p);  beta: pair q = (label q-1, blank q)
    int lab_off[K];
    unsigned skipmask = 0;
    const int pbase = rw * 32 * K + lane;
    if (compute || is_post) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int pp = pbase + k * 32;
            const int li = dir ? pp - 1 : pp;  // natural index of my label
            int off = c.slot_bytes - 16;       // sentinel words: emission log(0) -> state stays log(0)
            if (li >= 0 && li < L) {
                int l = tg[li];
                l = l < 0 ? 0 : (l >= V ? V - 1 : l);
                off = 4 * l;
                const int lo = dir ? li + 1 : li - 1;  // the label a skip transition comes from
                if (lo >= 0 && lo < L) {
                    int l2 = tg[lo];
                    l2 = l2 < 0 ? 0 : (l2 >= V ? V - 1 : l2);
                    if (l2 != l) skipmask |= 1u << k;
                }
            }
            lab_off[k] = off;
        }
    }

    // ---- label-sorted order of the label states (backward only; deterministic) ----
    // occ_start[c] .. occ_start[c+1] are the slots of label c; occ_pos[i] is the slot of natural label i.
    // The recursion warps publish label posteriors directly in this order, so that the gradient
    // warps sum contiguous runs.
    if (GRAD) {
        for (int cc = tid; cc < V; cc += blockDim.x) cursor[cc] = 0;
        __syncthreads();
        for (int i = tid; i < L; i += blockDim.x) {
            int l = tg[i];
            l = l < 0 ? 0 : (l >= V ? V - 1 : l);
            atomicAdd(&cursor[l], 1);
        }
        __syncthreads();
        if (warp == 0) {
            int run = 0;
            for (int c0 = 0; c0 < V; c0 += 32) {
                const int cc = c0 + lane;
                const int n = cc < V ? cursor[cc] : 0;
                int incl = n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int y = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += y;
                }
                if (cc < V) {
                    occ_start[cc] = run + incl - n;
                    cursor[cc] = run + incl - n;
                }
                run += __shfl_sync(FULL, incl, 31);
            }
            if (lane == 0) occ_start[V] = run;
            __syncwarp();
            for (int i0 = 0; i0 < L; i0 += 32) {
                const int i = i0 + lane;
                const unsigned act = __ballot_sync(FULL, i < L);
                if (i < L) {
                    int l = tg[i];
                    l = l < 0 ? 0 : (l >= V ? V - 1 : l);
                    const unsigned mm = __match_any_sync(act, l);
                    const int rank = __popc(mm & ((1u << lane) - 1u));
                    const int base = cursor[l];
                    occ_pos[i] = base + rank;  // natural label index -> slot in label-sorted order
                    __syncwarp(act);
                    if (rank == 0) cursor[l] = base + __popc(mm);
                }
                __syncwarp();
            }
            // compact list of the columns with posterior mass (labels of this utterance + blank), ascending
            int np = 0;
            for (int c0 = 0; c0 < V; c0 += 32) {
                const int cc = c0 + lane;
                const bool has = cc < V && (occ_start[cc + 1] > occ_start[cc] || cc == p.blank);
                const unsigned bal = __ballot_sync(FULL, has);
                __syncwarp();
                if (has) cursor[np + __popc(bal & ((1u << lane) - 1u))] = cc;
                np += __popc(bal);
            }
            if (lane == 0) occ_start[V + 1] = np;
        }
    }

    // ---- recursion state: virtual start row (forward) or the stored frontier (backward) ----
    float ab[K], al[K];
    double off_mine = 0.0;  // accumulated re-centring offset: true value = state + off_mine
    const int lab_pos = P_pad + 1 - dir + pbase;  // row position of my k=0 label state
    if (compute) {
        const float *fin = p.finals + ((int64_t)b * 2 + dir) * row_elems;
        // frontier written by the wavefront forward: per-warp offsets (see "offset tables"); continue from the
        // largest one, so that the states that matter start out small
        const int Kf = GRAD ? (int)fin[2 * P_pad + 4] : 0;
        const float *tab = p.tabs + (((int64_t)b * 2 + dir) * p.NCH + (p.NCH - 1)) * 16;
        float Dmax = 0.f;
        if (GRAD) {
            off_mine = *reinterpret_cast<const double *>(fin + 2 * P_pad + 2);
            if (Kf > 0) {
                const int nsl = L / (32 * (Kf > 0 ? Kf : 1)) + 1;
                Dmax = tab[0];
                for (int w = 1; w < nsl; ++w) Dmax = fmaxf(Dmax, tab[w]);
                off_mine += (double)Dmax;
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int pp = pbase + k * 32;
            if (GRAD) {
                const int li = dir ? pp - 1 : pp;
                const int cb_ = dir ? L - pp : pp, cl_ = dir ? L - li - 1 : li;   // chain elements of my two states
                const int span_f = 32 * (Kf > 0 ? Kf : 1);
                const float tb_ = Kf > 0 && pp <= L ? tab[cb_ / span_f] - Dmax : 0.f;
                const float tl_ = Kf > 0 && li >= 0 && li < L ? tab[cl_ / span_f] - Dmax : 0.f;
                ab[k] = pp <= L ? fin[pp] + tb_ : kNeg;                                        // pairs beyond L
                al[k] = lab_off[k] == c.slot_bytes - 16 ? kNeg : fin[lab_pos + k * 32] + tl_;  // states beyond 2L+1
            } else {
                ab[k] = pp == (dir ? L : 0) ? 0.f : kNeg;
                al[k] = kNeg;
            }
        }
        if (lane == (dir ? 0 : 31)) xchg[1 + warp] = dir ? al[0] : al[K - 1];
    }
    __syncthreads();  // mbarrier init, CSR, xchg visible; last CTA-wide barrier

    constexpr int C = CH;
    const int NST = c.stages;
    const int nchunks = (nsteps + C - 1) / C;
    const int64_t step_elems = (int64_t)dt * p.st;
    const float *first_row = lp_b + (int64_t)t_first * p.st;

    if (producer) {
        // ================= producer warps: bulk-copy issue, blocking mbarrier waits only ==========
        // warp W streams the emission rows; in the backward call warp W+1 streams the other
        // direction's stored lattice rows.  try_wait suspends the warp in hardware: no polling.
        if (warp == W) {
            RingProducer prod;
            prod.src = first_row;
            prod.step_elems = step_elems;
            prod.stage = 0;
            prod.remaining = nsteps;
            int round = 0;
            for (int n = 0; n < nchunks; ++n) {
                if (round > 0) mbar_wait(&em_empty[prod.stage], (uint32_t)((round - 1) & 1));
                const int stg = prod.stage;
                if (lane == 0) ring_issue_next(ring, prod);
                prod.stage = __shfl_sync(FULL, prod.stage, 0);
                if (prod.stage <= stg) ++round;
            }
        } else {
            const int Co = c.or_chunk, No = c.or_stages;
            const int or_nchunks = (nsteps + Co - 1) / Co;
            int or_stage = 0, or_round = 0, or_left = nsteps;
            const float *or_src = p.rows + ((int64_t)(p.slot ? p.slot[b] : b) * p.T + t_first) * row_elems;
            const int64_t or_step = (int64_t)dt * row_elems;
            for (int n = 0; n < or_nchunks; ++n) {
                if (or_round > 0) mbar_wait(&or_empty[or_stage], (uint32_t)((or_round - 1) & 1));
                const int cnt = or_left < Co ? or_left : Co;
                if (lane == 0) {
                    mbar_arrive_expect_tx(&or_full[or_stage], (uint32_t)(cnt * row_bytes));
                    unsigned char *dst = or_slots + (size_t)or_stage * Co * row_bytes;
                    for (int f = 0; f < cnt; ++f, dst += row_bytes)
                        bulk_g2s(dst, or_src + (int64_t)f * or_step, (uint32_t)row_bytes, &or_full[or_stage]);
                }
                or_src += (int64_t)cnt * or_step;
                or_left -= cnt;
                if (++or_stage == No) { or_stage = 0; ++or_round; }
            }
        }
        return;
    }

    // ================= recursion and gradient warps ==========================================
    const int nbar = wl * 32;  // the per-frame barrier is among the live recursion warps only
    const int slot_bytes = c.slot_bytes;
    const unsigned char *em_base = ring.slots;
    const int Co = c.or_chunk, No = c.or_stages, or_nslots = Co * No;
    const int64_t st_step = (int64_t)dt * row_elems;
    const int64_t grow_step = (int64_t)dt * p.gst;
    const int blank_off = 4 * p.blank;
    const unsigned a15_0 = (unsigned)(reinterpret_cast<uintptr_t>(first_row) & 15);
    const unsigned a15_step = (unsigned)((step_elems * 4) & 15);  // CH * a15_step % 16 == 0
    unsigned skip_m[K], bl_m[K];  // bl_m: my k-th blank state exists (pair <= L)
    int wl_off[K];  // byte offset of my label posterior in the label-sorted buffer (backward)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        skip_m[k] = (skipmask >> k) & 1u ? 0xffffffffu : 0u;
        bl_m[k] = pbase + k * 32 <= L ? 0xffffffffu : 0u;
        wl_off[k] = 4 * (WL - 1);  // dump slot for the states beyond 2L+1
        if (GRAD && (compute || is_post)) {
            const int li = dir ? pbase + k * 32 - 1 : pbase + k * 32;
            if (li >= 0 && li < L) wl_off[k] = 4 * occ_pos[li];
        }
    }
    auto sel = [](unsigned m, float a, float bb) {  // m ? a : bb, one LOP3
        return __int_as_float((__float_as_int(a) & m) | (__float_as_int(bb) & ~m));
    };

    if (dead) {
        // nothing: straight to the common tail (frames beyond the utterance)
    } else if (compute) {
        // ---------------- recursion warps: chunk-unrolled time loop ----------------
        auto run = [&](auto dir_tag) {
            constexpr int DIR = decltype(dir_tag)::value;
            const unsigned seam_m = lane == (DIR ? 31 : 0) ? 0xffffffffu : 0u;
            const bool is_out = lane == (DIR ? 0 : 31);
            const int nb_lane = DIR ? (lane + 1) & 31 : (lane + 31) & 31;
            const float *x_in = xchg + 1 + warp + (DIR ? 1 : -1);
            float *x_out = xchg + 1 + warp;
            // row pointers: pair (pbase + 32k) -> blank at [pbase+32k], label at [P_pad+1-DIR+pbase+32k]
            const bool save = !GRAD && p.rows != nullptr;
            const int lab_delta = P_pad + 1 - DIR;
            float *sb = save ? p.rows + ((int64_t)(p.slot ? p.slot[b] : b) * p.T + t_first) * row_elems + pbase : nullptr;  // my blank states
            float *sl = sb + lab_delta;                                                                // my label states
            double *soff = reinterpret_cast<double *>(sb - pbase + 2 * P_pad + 2);                    // the row's offset
            const unsigned char *or_row = or_slots;
            int or_slot = 0, or_left = 0, ostage = 0, ophase = 0;
            unsigned char *wl_bytes = reinterpret_cast<unsigned char *>(wlab);
            int em_stage = 0, em_phase = 0;
            const unsigned char *em_chunk = em_base;
            int remaining = nsteps;
            bool first_chunk = true;
            int chunk_idx = 0;
            float xfix = 0.f;
            double base_d = 0.0;
            int pbuf = 0;
            // raw logits: the row normaliser -log2 sum exp of frame f of the chunk sits in lane f, fetched one
            // chunk ahead; emissions are x * log2(e) + zl
            const float *z_ptr = LOGITS ? p.zl + (int64_t)t_first * p.B + b : nullptr;
            const int64_t z_step = (int64_t)dt * p.B;
            auto z_fetch = [&](int step0) -> float {
                const int sidx = step0 + lane;
                return (LOGITS && lane < CH && sidx < nsteps) ? __ldg(z_ptr + (int64_t)sidx * z_step) : 0.f;
            };
            float zv = 0.f, zv_next = z_fetch(0);
            int step0 = 0;
            float nan_acc = 0.f;
            // one frame; `n` is the number of frames of the current chunk (a constant CH on the fast path)
            auto frame = [&](const int f, const int n) {
                // the bulk copy lands the row (addr & 15) bytes into its slot; raw natural-log values
                const unsigned char *row = em_chunk + f * slot_bytes + ((a15_0 + f * a15_step) & 15u);
                const float zl = LOGITS ? __shfl_sync(FULL, zv, f) : 0.f;
                const float eb_s = LOGITS ? fmaf(*reinterpret_cast<const float *>(row + blank_off), kLog2e, zl)
                                          : *reinterpret_cast<const float *>(row + blank_off) * kLog2e;
                const float eb2 = fmaxf(eb_s, kNeg);
                // sticky NaN detector on the blank column (fmaxf drops a NaN operand): a diverged model emits whole NaN
                // rows, and one add per frame costs no registers (summing the label columns as well cost 17 registers
                // and a resident CTA per SM in the many-CTA shapes)
                if (!GRAD) nan_acc += eb_s;
                float el2[K];
#pragma unroll
                for (int k = 0; k < K; ++k)
                    el2[k] = fmaxf(LOGITS ? fmaf(*reinterpret_cast<const float *>(row + lab_off[k]), kLog2e, zl)
                                          : *reinterpret_cast<const float *>(row + lab_off[k]) * kLog2e, kNeg);
                float xin = x_in[(f & 1) * 18];
                if (f == 0) xin -= xfix;
                float r[K];
#pragma unroll
                for (int k = 0; k < K; ++k) r[k] = __shfl_sync(FULL, al[k], nb_lane);
                float pre_b[K], pre_l[K];  // the states before the emission is added (what a posterior needs)
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float seamv = DIR ? (k == K - 1 ? xin : r[k < K - 1 ? k + 1 : k])
                                            : (k == 0 ? xin : r[k > 0 ? k - 1 : 0]);
                    const float carry = sel(seam_m, seamv, r[k]);
                    const float A = lse2(ab[k], carry);
                    const float oth = sel(skip_m[k], A, ab[k]);
                    pre_b[k] = A;
                    pre_l[k] = lse2(al[k], oth);
                    ab[k] = A + eb2;
                    al[k] = pre_l[k] + el2[k];
                }
                if (is_out) x_out[((f & 1) ^ 1) * 18] = DIR ? al[0] : al[K - 1];
                if (f == CH - 1) {  // publish the row maximum for the next chunk's re-centring
                    float mx = kNeg;
#pragma unroll
                    for (int k = 0; k < K; ++k) mx = fmaxf(mx, fmaxf(ab[k], al[k]));
                    mx = warp_max(mx);
                    if (lane == 0) wmax[warp] = mx;
                }
                if (!GRAD) {
                    if (save) {
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            sb[k * 32] = ab[k];
                            sl[k * 32] = al[k];
                        }
                        if (tid == 0) *soff = off_mine;
                        sb += st_step;
                        sl += st_step;
                        soff += st_step / 2;  // row_elems is even
                    }
                } else if (split) {
                    // hand my states of this frame to my posterior warp (state ring, buffer of this chunk)
                    float *sr = sring + (size_t)(pbuf + f) * 2 * P_pad + pbase;
#pragma unroll
                    for (int k = 0; k < K; ++k) {   // alpha + beta - lp = (state before its emission) + other direction
                        sr[k * 32] = pre_b[k];
                        sr[P_pad + k * 32] = pre_l[k];
                    }
                    if (f == n - 1) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&st_full[warp * 2 + (chunk_idx & 1)]);
                    }
                } else {
                    // posteriors of my states at this frame: 2^(alpha + beta - lp - log2 P)
                    if (or_left == 0) {
                        mbar_wait(&or_full[ostage], (uint32_t)ophase);
                        or_left = Co;
                    }
                    const float *orow = reinterpret_cast<const float *>(or_row) + pbase;
                    const double ooff = *reinterpret_cast<const double *>(or_row + (2 * P_pad + 2) * 4);
                    const float bracket = (float)(base_d + ooff);
                    const float cb = bracket - eb2;
                    unsigned char *wl = wl_bytes + (pbuf + f) * (WL * 4);
                    float sbl = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        // (the other direction's row holds nothing for the pairs beyond L)
                        sbl += ex2_approx(ab[k] + sel(bl_m[k], orow[k * 32], kNeg) + cb);
                        *reinterpret_cast<float *>(wl + wl_off[k]) =
                            ex2_approx(al[k] + orow[lab_delta + k * 32] + (bracket - el2[k]));
                    }
                    const unsigned fx = __float2uint_rn(fminf(sbl, 3.5f) * 1073741824.0f);
                    const unsigned tot = __reduce_add_sync(FULL, fx);
                    __syncwarp();
                    if (lane == 0) {
                        atomicAdd(&blank_acc[pbuf + f], tot);
                        mbar_arrive(&post_full[pbuf + f]);  // release: this warp's posteriors of the frame
                    }
                    or_row += row_bytes;
                    if (++or_slot == or_nslots) { or_slot = 0; or_row = or_slots; }
                    if (--or_left == 0 || (f == n - 1 && remaining == n)) {  // stage done / last frame
                        or_left = 0;
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&or_empty[ostage]);
                        if (++ostage == No) { ostage = 0; ophase ^= 1; }
                    }
                }
                if (f == n - 1) {  // release the emission stage before the barrier
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&em_empty[em_stage]);
                }
                named_bar_sync(1, nbar);
            };
            while (remaining > 0) {
                const int n = remaining < CH ? remaining : CH;
                if (LOGITS) {
                    zv = zv_next;
                    step0 += CH;
                    zv_next = z_fetch(step0);
                }
                mbar_wait(&em_full[em_stage], (uint32_t)em_phase);
                xfix = 0.f;  // correction for the seam value written before the re-centring
                if (!first_chunk) {
                    // re-centre on the row maximum published at the end of the previous chunk
                    float mx = wmax[0];
                    for (int w = 1; w < wl; ++w) mx = fmaxf(mx, wmax[w]);
                    mx = floorf(mx);  // integer amounts: the subtraction is exact, every stored offset an integer
                    if (mx > kNegTest) {
#pragma unroll
                        for (int k = 0; k < K; ++k) { ab[k] -= mx; al[k] -= mx; }
                        xfix = mx;
                        off_mine += (double)mx;
                    }
                }
                first_chunk = false;
                base_d = off_mine + nll2;
                if (GRAD && split) {
                    // my posterior warp must be done with this state buffer; it also needs this chunk's base
                    if (chunk_idx >= 2) mbar_wait(&st_empty[warp * 2 + (chunk_idx & 1)], (uint32_t)(((chunk_idx >> 1) - 1) & 1));
                    if (lane == 0) sbase[warp * 2 + (chunk_idx & 1)] = base_d;
                } else if (GRAD && chunk_idx >= 2) {  // the gradient warps must be done with this posterior buffer
                    mbar_wait(&post_empty[chunk_idx & 1], (uint32_t)(((chunk_idx >> 1) - 1) & 1));
                }
                pbuf = (chunk_idx & 1) * CH;
                if (n == CH) {  // fast path: full chunk, every per-frame test folds at compile time
#pragma unroll
                    for (int f = 0; f < CH; ++f) frame(f, CH);
                } else {
#pragma unroll 1
                    for (int f = 0; f < n; ++f) frame(f, n);
                }
                remaining -= n;
                ++chunk_idx;
                em_chunk += CH * slot_bytes;
                if (++em_stage == NST) { em_stage = 0; em_phase ^= 1; em_chunk = em_base; }
            }
            if (!GRAD && __any_sync(FULL, nan_acc != nan_acc) && lane == 0) p.nan_flag[b] = 1;
        };
        if (dir) run(std::integral_constant<int, 1>{}); else run(std::integral_constant<int, 0>{});
    } else if (GRAD && is_post) {
        // ---------------- posterior warps: one chunk behind their recursion warp ----------------
        // posterior of a state = 2^(alpha + beta - lp - log2 P): my recursion warp's states (state ring) times
        // the other direction's stored row (row ring), label posteriors into the label-sorted posterior ring,
        // blank posteriors summed in fixed point.  Everything here is off the serial chain of the recursion.
        const int lab_delta = P_pad + 1 - dir;
        // offset-table slots of my states in the OTHER direction's rows (see "offset tables"; Kf = 0: none)
        const int Kf = (int)(p.finals + ((int64_t)b * 2 + dir) * row_elems)[2 * P_pad + 4];
        const bool tab_on = Kf > 0;
        const float *otabs = p.tabs + ((int64_t)b * 2 + (1 - dir)) * p.NCH * 16;
        int tsb[K], tsl[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int pp = pbase + k * 32, li = dir ? pp - 1 : pp;
            const int cbo = dir ? pp : L - pp, clo = dir ? li : L - li - 1;
            tsb[k] = (tab_on && pp <= L) ? cbo / (32 * Kf) : 0;
            tsl[k] = (tab_on && li >= 0 && li < L) ? clo / (32 * Kf) : 0;
        }
        const unsigned char *or_row = or_slots;
        int or_slot = 0, or_left = 0, ostage = 0, ophase = 0;
        unsigned char *wl_bytes = reinterpret_cast<unsigned char *>(wlab);
        int remaining = nsteps, chunk_idx = 0;
        while (remaining > 0) {
            const int n = remaining < CH ? remaining : CH;
            const int buf = chunk_idx & 1, pbuf = buf * CH;
            mbar_wait(&st_full[rw * 2 + buf], (uint32_t)((chunk_idx >> 1) & 1));   // the chunk's states are there
            if (chunk_idx >= 2)  // the gradient warps must be done with this posterior buffer
                mbar_wait(&post_empty[buf], (uint32_t)(((chunk_idx >> 1) - 1) & 1));
            // bracket = base + row offset (+ table entry).  Row offsets and table entries are integers below 2^24
            // (exact in fp32): only the base has a fraction, split off once per chunk -- no fp64 per frame or state.
            const double base_d = sbase[rw * 2 + buf], base_fl = floor(base_d);
            const float base_i = (float)base_fl, br_f = (float)(base_d - base_fl);
            // offset tables: the other direction computed my frames in DEcreasing step order, so this chunk sees at
            // most two of its chunks -- table A for the frames f <= fsw, table B = the one before it afterwards
            float ta_b[K], ta_l[K], tb_b[K], tb_l[K];
            int fsw = CH;
            {
                const int t0 = t_first + dt * (chunk_idx * CH);
                const int tau0o = dir ? t0 : Tb - 1 - t0;          // the other direction's forward step of frame 0
                fsw = tau0o % CH;
                const float *pa = otabs + (tau0o / CH) * 16, *pb = pa - (tau0o >= CH ? 16 : 0);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    ta_b[k] = tab_on ? __ldg(pa + tsb[k]) : 0.f;
                    ta_l[k] = tab_on ? __ldg(pa + tsl[k]) : 0.f;
                    tb_b[k] = tab_on ? __ldg(pb + tsb[k]) : 0.f;
                    tb_l[k] = tab_on ? __ldg(pb + tsl[k]) : 0.f;
                }
            }
            const bool chunked_rows = Co == CH;   // the row ring's stages are this loop's chunks: wait / release once
            if (chunked_rows) mbar_wait(&or_full[ostage], (uint32_t)ophase);
            auto pframe = [&](const int f) {
                if (!chunked_rows && or_left == 0) {
                    mbar_wait(&or_full[ostage], (uint32_t)ophase);
                    or_left = Co;
                }
                const float *orow = reinterpret_cast<const float *>(or_row) + pbase;
                const float br_i = base_i + (float)*reinterpret_cast<const double *>(or_row + (2 * P_pad + 2) * 4);
                const bool useA = f <= fsw;
                const float *sr = sring + (size_t)(pbuf + f) * 2 * P_pad + pbase;
                unsigned char *wl = wl_bytes + (pbuf + f) * (WL * 4);
                float sbl = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) {   // (the state ring holds the states before their emission: no lp here)
                    const float br_b = (br_i + (useA ? ta_b[k] : tb_b[k])) + br_f;
                    const float br_l = (br_i + (useA ? ta_l[k] : tb_l[k])) + br_f;
                    sbl += ex2_approx(sr[k * 32] + sel(bl_m[k], orow[k * 32], kNeg) + br_b);
                    *reinterpret_cast<float *>(wl + wl_off[k]) =
                        ex2_approx(sr[P_pad + k * 32] + orow[lab_delta + k * 32] + br_l);
                }
                const unsigned fx = __float2uint_rn(fminf(sbl, 3.5f) * 1073741824.0f);
                const unsigned tot = __reduce_add_sync(FULL, fx);
                __syncwarp();
                if (lane == 0) {
                    atomicAdd(&blank_acc[pbuf + f], tot);
                    mbar_arrive(&post_full[pbuf + f]);  // release: this warp's posteriors of the frame
                }
                or_row += row_bytes;
                if (!chunked_rows) {
                    if (++or_slot == or_nslots) { or_slot = 0; or_row = or_slots; }
                    if (--or_left == 0 || (f == n - 1 && remaining == n)) {  // stage done / last frame
                        or_left = 0;
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&or_empty[ostage]);
                        if (++ostage == No) { ostage = 0; ophase ^= 1; }
                    }
                }
            };
            // (a compact loop on purpose: three big code regions -- recursion, posterior, gradient -- share the
            //  instruction cache of the SM; unrolling this one made the backward 15 % slower)
#pragma unroll 1
            for (int f = 0; f < n; ++f) pframe(f);
            if (chunked_rows) {   // release the row-ring stage; a partial last chunk still owns a whole stage
                __syncwarp();
                if (lane == 0) mbar_arrive(&or_empty[ostage]);
                if (++ostage == No) { ostage = 0; ophase ^= 1; or_row = or_slots; }
                else or_row = or_slots + (size_t)ostage * Co * row_bytes;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&st_empty[rw * 2 + buf]);
            remaining -= n;
            ++chunk_idx;
        }
    } else if (GRAD) {
        // ---------------- gradient warps: consume the posterior ring, mbarriers only ----------------
        // Each gradient warp takes whole frames (frame f of a chunk goes to warp f mod G), so the
        // per-frame latency (barrier probe, run sums, exp, store) overlaps across warps.
        const int gwarp = warp - (W + NPROD), G = c.G;
        // G <= CH: warp g takes the frames g, g+G, ... of a chunk.  G = 2 CH: warps g and g+CH share frame g and
        // take the columns [0, V/2) and [V/2, V) (split at a multiple of 128 columns).
        const int GF = G > CH ? CH : G, gf = gwarp % GF, gpart = gwarp / GF;
        const int c_split = G > CH ? ((V / 2) & ~127) : V;
        const int c_lo = gpart ? c_split : 0, c_hi = (G > CH && gpart == 0) ? c_split : V;
        float *grow_chunk = p.grad + (int64_t)t_first * p.gst + (int64_t)b * p.gsb;
        const unsigned char *em_chunk = em_base;
        int em_stage = 0, em_phase = 0;
        // 128-bit path when every emission row and every gradient row is 16-byte aligned
        const bool vec = a15_0 == 0 && a15_step == 0 && (V & 3) == 0 && ((p.gst | p.gsb) & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(p.grad) & 15) == 0;
        // scalar path: bit j of `present` = my j-th column (lane + 32 j) carries posterior mass
        unsigned present = 0;
        if (!vec) {
            int j = 0;
            for (int cc = lane; cc < V; cc += 32, ++j)
                if (j < 32 && (occ_start[cc + 1] > occ_start[cc] || cc == p.blank)) present |= 1u << j;
        }
        auto label_mass = [&](int cc, const float *w, int slot) {  // posterior mass of column cc at this frame
            float rsum = 0.f;
            const int q1 = occ_start[cc + 1];
            // contiguous run (label-sorted order).  (Four partial sums were tried: runs are 1-3 long for most
            // labels and the extra instructions made the C2 backward 10 % slower.)
            for (int q = occ_start[cc]; q < q1; ++q) rsum += w[q];
            if (cc == p.blank) {
                rsum += (float)blank_acc[slot] * (1.0f / 1073741824.0f);
                blank_acc[slot] = 0u;
            }
            return rsum;
        };
        int remaining = nsteps, chunk_idx = 0;
        const float *z_ptr = LOGITS ? p.zl + (int64_t)t_first * p.B + b : nullptr;
        const int64_t z_step = (int64_t)dt * p.B;
        auto z_fetch = [&](int step0) -> float {
            const int sidx = step0 + lane;
            return (LOGITS && lane < CH && sidx < nsteps) ? __ldg(z_ptr + (int64_t)sidx * z_step) : 0.f;
        };
        float zv = 0.f, zv_next = z_fetch(0);
        while (remaining > 0) {
            const int n = remaining < CH ? remaining : CH;
            if (LOGITS) {
                zv = zv_next;
                zv_next = z_fetch((chunk_idx + 1) * CH);
            }
            mbar_wait(&em_full[em_stage], (uint32_t)em_phase);
            const int pbuf = (chunk_idx & 1) * CH;
            for (int f = gf; f < n; f += GF) {
                const int slot = pbuf + f;
                const float zl = LOGITS ? __shfl_sync(FULL, zv, f) : 0.f;
                mbar_wait(&post_full[slot], (uint32_t)((chunk_idx >> 1) & 1));
                const float *w = wlab + slot * WL;
                const unsigned char *rowb = em_chunk + f * slot_bytes + ((a15_0 + f * a15_step) & 15u);
                float *grow = grow_chunk + (int64_t)f * grow_step;
                if (vec) {
                    // dense pass: exp(lp) for every column, 128-bit loads and stores, no divergence
                    const float4 *row4 = reinterpret_cast<const float4 *>(rowb);
                    float4 *g4 = reinterpret_cast<float4 *>(grow);
                    for (int it = (c_lo >> 2) + lane; it < (c_hi >> 2); it += 32) {
                        const float4 x = row4[it];
                        g4[it] = make_float4(ex2_approx(fmaf(x.x, kLog2e, zl)) * gs, ex2_approx(fmaf(x.y, kLog2e, zl)) * gs,
                                             ex2_approx(fmaf(x.z, kLog2e, zl)) * gs, ex2_approx(fmaf(x.w, kLog2e, zl)) * gs);
                    }
                    __syncwarp();  // orders the dense stores before the overwrites below (same warp)
                    // sparse pass: the few columns that carry posterior mass, one per lane
                    const float *row = reinterpret_cast<const float *>(rowb);
                    const int np = occ_start[V + 1];
                    for (int i = lane; i < np; i += 32) {
                        const int cc = cursor[i];
                        if (cc >= c_lo && cc < c_hi)
                            grow[cc] = (ex2_approx(fmaf(row[cc], kLog2e, zl)) - label_mass(cc, w, slot)) * gs;
                    }
                } else {
                    const float *row = reinterpret_cast<const float *>(rowb);
                    int j = 0;
                    for (int cc = lane; cc < V; cc += 32, ++j) {
                        if (cc < c_lo || cc >= c_hi) continue;
                        float val = ex2_approx(fmaf(row[cc], kLog2e, zl));
                        if (j >= 32 || ((present >> j) & 1u)) val -= label_mass(cc, w, slot);
                        grow[cc] = val * gs;
                    }
                }
            }
            // release the emission stage and the posterior buffer
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&em_empty[em_stage]);
                mbar_arrive(&post_empty[chunk_idx & 1]);
            }
            remaining -= n;
            ++chunk_idx;
            grow_chunk += (int64_t)CH * grow_step;
            em_chunk += CH * slot_bytes;
            if (++em_stage == NST) { em_stage = 0; em_phase ^= 1; em_chunk = em_base; }
        }
    }

    if (!GRAD) {
        // frontier row for the join kernel / the backward call
        if (compute) {
            float *fin = p.finals + ((int64_t)b * 2 + dir) * row_elems + pbase;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                fin[k * 32] = ab[k];
                fin[P_pad + 1 - dir + k * 32] = al[k];
            }
            if (tid == 0) {
                *reinterpret_cast<double *>(fin + 2 * P_pad + 2) = off_mine;
                fin[2 * P_pad + 4] = 0.f;   // Kf = 0: barrier format, no offset tables
            }
        }
    } else if (dir == 0) {
        const int nthr = n_consumers * 32, me = warp < W ? tid : tid - 32 * NPROD;  // every warp but the producers
        for (int t = Tb; t < (int)p.T; ++t) {  // frames beyond the utterance: exact zeros
            float *g = p.grad + (int64_t)t * p.gst + (int64_t)b * p.gsb;
            for (int cc = me; cc < V; cc += nthr) g[cc] = 0.f;
        }
    }
}

// ------------------------------------------------------------------------------ wavefront forward kernel
// The forward launch (half lattices + frontier rows) without the per-frame barrier: chain element c -- the pair
// (blank, label) number c in recursion order, i.e. pair c for alpha and pair L-c for beta -- depends on elements
// c and c-1 of the previous frame only, so recursion warp w needs ONE value per frame from warp w-1.  Warps run
// skewed in time: warp w-1 streams its last label state ("seam") chunk by chunk into a small shared-memory ring
// whose words validate themselves (a NaN pattern = "not there yet", no flag / fence / barrier), warp w follows
// a chunk or two behind.  Lanes own K CONSECUTIVE chain elements (one shuffle per frame instead of K), the frame
// body is branch-free and unrolled over the chunk, stores to shared memory happen once per chunk.
//   Re-centring is per warp and by INTEGER amounts (exact in fp32; offsets add up exactly in int32): every chunk a
// warp subtracts floor(max of its states).  With the seam values travel the sender's offset and the offset of
// the chain's first warp; the latter is the row offset stored with every saved row, and every warp records
// off_w - off_row per chunk in the offset tables (see "offset tables"): the stored states stay small wherever the
// probability mass sits, the posterior warps / the join kernel add the integers back exactly.
struct FwdWaveSmem {
    int em_full, em_empty, seam, ring, total;
};
__host__ __device__ __forceinline__ FwdWaveSmem fwd_wave_smem(int W, int stages, int chunk, int slot_bytes) {
    FwdWaveSmem m;
    m.em_full = 0;
    m.em_empty = 8 * stages;
    m.seam = 16 * stages;                                        // [W+1][stages+1][chunk+2] words
    m.ring = (m.seam + 4 * (W + 1) * (stages + 1) * (chunk + 2) + 127) & ~127;
    m.total = m.ring + stages * chunk * slot_bytes;
    return m;
}
constexpr uint32_t kSeamEmptyBits = 0xffffffffu;
__device__ __forceinline__ float ld_volatile_shared_f32(const float *p) {
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}

struct FwdWaveCfg {
    int K, W, stages, chunk, slot_bytes;
};

template <int K, int CH, bool LOGITS, bool SAVE>
__global__ void __launch_bounds__(K == 1 ? 544 : 288, 1) ctc_forward_wave_kernel(const CtcParams p, const FwdWaveCfg wc) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int b = utterance_of_cta(p), dir = blockIdx.y;
    if (b < 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const int W = wc.W;
    const float EMPTY = __uint_as_float(kSeamEmptyBits);

    int Tb = p.in_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.T ? (int)p.T : Tb);
    int L = p.tgt_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    const int m = Tb >> 1;
    const int nsteps = dir ? Tb - m : m;            // frames this direction owns in forward()
    const int P_pad = p.cfg.P_pad, row_elems = p.cfg.row_elems;
    const int V = p.V;
    const float *lp_b = p.lp + (int64_t)b * p.sb;
    const int32_t *tg = p.targets + p.tgt_off[b];
    const int t_first = dir ? Tb - 1 : 0;
    const int dt = dir ? -1 : 1;

    int wlive = L / (32 * K) + 1;                   // chain elements 0..L
    wlive = wlive > W ? W : wlive;
    const bool compute = warp < wlive;

    const int NST = wc.stages, NSLOT = NST + 1, slot_bytes = wc.slot_bytes, SW = CH + 2;
    const FwdWaveSmem lay = fwd_wave_smem(W, NST, CH, slot_bytes);
    uint64_t *em_full = reinterpret_cast<uint64_t *>(smem + lay.em_full);
    uint64_t *em_empty = reinterpret_cast<uint64_t *>(smem + lay.em_empty);
    float *seam = reinterpret_cast<float *>(smem + lay.seam);  // [W+1][NSLOT][SW]
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
    for (int i = tid; i < (W + 1) * NSLOT * SW; i += blockDim.x) seam[i] = EMPTY;
    // sentinel emission log(0) for the chain elements beyond L: the last 16 bytes of a slot are never written
    for (int i = tid; i < NST * CH * 4; i += blockDim.x)
        *reinterpret_cast<float *>(ring.slots + (size_t)(i >> 2) * slot_bytes + slot_bytes - 16 + (i & 3) * 4) = kNeg;
    __syncthreads();  // the only CTA-wide barrier
    if (warp >= wlive && warp != W) return;

    const int nchunks = (nsteps + CH - 1) / CH;
    const int64_t step_elems = (int64_t)dt * p.st;
    const float *first_row = lp_b + (int64_t)t_first * p.st;
    if (!compute) {
        // ================= producer warp =================
        RingProducer prod;
        prod.src = first_row;
        prod.step_elems = step_elems;
        prod.stage = 0;
        prod.remaining = nsteps;
        int round = 0;
        for (int n = 0; n < nchunks; ++n) {
            if (round > 0) mbar_wait(&em_empty[prod.stage], (uint32_t)((round - 1) & 1));
            const int stg = prod.stage;
            if (lane == 0) ring_issue_next(ring, prod);
            prod.stage = __shfl_sync(FULL, prod.stage, 0);
            if (prod.stage <= stg) ++round;
        }
        return;
    }

    // ================= recursion warps =================
    // chain element c: alpha -> (blank c, label c); beta -> (blank L-c, label L-c-1)
    const int cbase = warp * 32 * K + lane * K;
    int lab_off[K], bpos[K], lpos[K];
    unsigned skip_m[K];
    float ab[K], al[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int c = cbase + k;
        const int li = dir ? L - c - 1 : c;          // natural index of my label
        int off = slot_bytes - 16;                   // sentinel word: emission log(0)
        skip_m[k] = 0u;
        if (c <= L && li >= 0 && li < L) {
            int l = tg[li];
            l = l < 0 ? 0 : (l >= V ? V - 1 : l);
            off = 4 * l;
            const int lo = dir ? li + 1 : li - 1;    // the label a skip transition comes from
            if (lo >= 0 && lo < L) {
                int l2 = tg[lo];
                l2 = l2 < 0 ? 0 : (l2 >= V ? V - 1 : l2);
                if (l2 != l) skip_m[k] = 0xffffffffu;
            }
        }
        lab_off[k] = off;
        // row positions; states that do not exist go to the spare word [P_pad], which nobody reads for real
        bpos[k] = c <= L ? (dir ? L - c : c) : P_pad;
        lpos[k] = (c <= L && li >= 0 && li < L) ? P_pad + 1 + li : P_pad;
        ab[k] = c == 0 ? 0.f : kNeg;                 // virtual start row
        al[k] = kNeg;
    }
    const bool first = warp == 0;
    const bool dn_live = (warp + 1) * 32 * K <= L;   // a live downstream warp reads my seam
    const float *sv_in = seam + (warp - 1) * NSLOT * SW;
    float *sv_out = seam + (dn_live ? warp : W) * NSLOT * SW;
    const bool is31 = lane == 31;
    const unsigned seam_m = lane == 0 ? 0xffffffffu : 0u;
    auto sel = [](unsigned mm, float a, float bb) {
        return __int_as_float((__float_as_int(a) & mm) | (__float_as_int(bb) & ~mm));
    };
    const int blank_off = 4 * p.blank;
    const unsigned a15_0 = (unsigned)(reinterpret_cast<uintptr_t>(first_row) & 15);
    const unsigned a15_step = (unsigned)((step_elems * 4) & 15);
    float *row_out = SAVE ? p.rows + ((int64_t)(p.slot ? p.slot[b] : b) * p.T + t_first) * row_elems : nullptr;
    const int64_t row_step = (int64_t)dt * row_elems;
    const unsigned char *em_base = ring.slots, *em_chunk = em_base;
    int em_stage = 0, em_phase = 0, remaining = nsteps, sslot = 0, step0 = 0;
    int off_me = 0, off_row = 0;                     // integer log2 offsets: true value = state + off
    bool has_mass = false;
    int chunk_no = 0;
    float *tabs_bd = p.tabs + ((int64_t)b * 2 + dir) * p.NCH * 16;
    float ds = 0.f;                                  // off_me - off_row: conversion to the stored row's offset
    double off_row_d = 0.0;
    const bool inlane = lane < SW;
    auto fetch_seam = [&](int slot) -> float {
        return (!first && inlane) ? ld_volatile_shared_f32(sv_in + slot * SW + lane) : kNeg;
    };
    float sv_pre = fetch_seam(0);
    const float *z_ptr = LOGITS ? p.zl + (int64_t)t_first * p.B + b : nullptr;
    const int64_t z_step = (int64_t)dt * p.B;
    auto z_fetch = [&](int s0) -> float {
        const int sidx = s0 + lane;
        return (LOGITS && lane < CH && sidx < nsteps) ? __ldg(z_ptr + (int64_t)sidx * z_step) : 0.f;
    };
    float zv = 0.f, zv_next = z_fetch(0);

    float sout[CH];
    float nan_acc = 0.f;
    auto frame = [&](const int f, const float sv, auto direct_tag) {
        constexpr bool DIRECT = decltype(direct_tag)::value;
        const unsigned char *row = em_chunk + f * slot_bytes + ((a15_0 + f * a15_step) & 15u);
        const float zl = LOGITS ? __shfl_sync(FULL, zv, f) : 0.f;
        const float xb = *reinterpret_cast<const float *>(row + blank_off);
        const float eb_s = LOGITS ? fmaf(xb, kLog2e, zl) : xb * kLog2e;
        const float eb2 = fmaxf(eb_s, kNeg);
        nan_acc += eb_s;   // sticky NaN detector on the blank column (fmaxf drops a NaN operand)
        float el2[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float x = *reinterpret_cast<const float *>(row + lab_off[k]);
            el2[k] = fmaxf(LOGITS ? fmaf(x, kLog2e, zl) : x * kLog2e, kNeg);
        }
        const float xin = __shfl_sync(FULL, sv, f);
        if (DIRECT) {
            if (is31) sv_out[sslot * SW + f] = al[K - 1];
        } else {
            sout[f] = al[K - 1];
        }
        const float rk = __shfl_sync(FULL, al[K - 1], (lane + 31) & 31);
        float carry = sel(seam_m, xin, rk);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            // One maximum PER SUM (a common maximum over the three terms flushes a whole sum to zero when the third
            // term dominates by 2^126: with tight alignments, T_b ~ L_b + repeats, and peaky emissions exactly those
            // states carry the likelihood).  The label sum takes the blank sum as (M_a, s_a) instead of waiting for
            // its logarithm, so the dependent chain per frame stays one log-sum-exp long: 4 EX2 + 2 LG2 per pair.
            const float M_a = fmaxf(ab[k], carry);
            const float o = sel(skip_m[k], M_a, ab[k]);
            const float M_l = fmaxf(al[k], o);
            const float e_b = ex2_approx(ab[k] - M_a), e_c = ex2_approx(carry - M_a);
            const float e_l = ex2_approx(al[k] - M_l), w = ex2_approx(o - M_l);
            const float s_a = e_b + e_c;                                  // in [1, 2]
            const float s_l = fmaf(sel(skip_m[k], s_a, 1.0f), w, e_l);    // in [1, 3]
            carry = al[k];
            al[k] = (M_l + lg2_approx(s_l)) + el2[k];
            ab[k] = fmaxf(M_a + lg2_approx(s_a), kNeg) + eb2;
        }
        if (SAVE) {
            float *r = row_out + (int64_t)f * row_step;
#pragma unroll
            for (int k = 0; k < K; ++k) {   // relative to row offset + my table entry (see "offset tables")
                r[bpos[k]] = ab[k];
                r[lpos[k]] = al[k];
            }
            if (first && lane == 0) *reinterpret_cast<double *>(r + 2 * P_pad + 2) = off_row_d;
        }
    };

    while (remaining > 0) {
        const int n = remaining < CH ? remaining : CH;
        // re-centre on floor(max of my states): exact, the offset stays an integer
        {
            float mx = kNeg;
#pragma unroll
            for (int k = 0; k < K; ++k) mx = fmaxf(mx, fmaxf(ab[k], al[k]));
            mx = floorf(warp_max(mx));
            has_mass = mx > kNegTest;
            if (has_mass && mx != 0.f) {
#pragma unroll
                for (int k = 0; k < K; ++k) { ab[k] -= mx; al[k] -= mx; }
                off_me += (int)mx;
            }
        }
        // incoming seam of this chunk (read early during the previous chunk; poll only if that was too soon):
        // lanes [0, n) the values, lane CH the sender's offset, lane CH+1 the row offset
        float sv = sv_pre;
        if (!first) {
            // The upstream warp is a warp of this CTA and never waits on anything downstream, so the poll always
            // ends; the watchdog (10 s of wall time) is a safety net that turns a bug into NaN likelihoods for the
            // call (join kernel) instead of a hang or a __trap() that would poison the caller's CUDA context.
            unsigned spins = 0;
            unsigned long long t0 = 0;
            bool gave_up = false;
            while (__any_sync(FULL, (lane < n || (lane >= CH && lane < SW)) && __float_as_uint(sv) == kSeamEmptyBits)) {
                sv = ld_volatile_shared_f32(sv_in + sslot * SW + (lane < SW ? lane : 0));
                if ((++spins & 0x3fffu) == 0) {
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    const bool stop = now - t0 > 10000000000ull || *reinterpret_cast<volatile int *>(p.abort_word) != 0;
                    if (__any_sync(FULL, stop)) { gave_up = true; break; }
                }
            }
            if (gave_up) {
                if (lane == 0) atomicExch(p.abort_word, 1);
                while (remaining > 0) {   // drain the emission ring: the producer waits for every live warp
                    mbar_wait(&em_full[em_stage], (uint32_t)em_phase);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&em_empty[em_stage]);
                    remaining -= remaining < CH ? remaining : CH;
                    if (++em_stage == NST) { em_stage = 0; em_phase ^= 1; }
                }
                return;
            }
            if (inlane) const_cast<float *>(sv_in)[sslot * SW + lane] = EMPTY;  // recycle the slot
            const int off_up = (int)__shfl_sync(FULL, sv, CH);
            off_row = (int)__shfl_sync(FULL, sv, CH + 1);
            // a warp the probability mass has not reached yet (all states log 0) follows the row offset, so that
            // its table entry stays 0 and the first mass enters it in the row's scale
            if (!has_mass) off_me = off_row;
            // the sender's values are relative to its offset: bring them to mine (lanes >= CH are not used)
            sv += (float)(off_up - off_me);
        } else {
            off_row = off_me;
        }
        ds = (float)(off_me - off_row);
        off_row_d = (double)off_row;
        if (SAVE && lane == 0) tabs_bd[chunk_no * 16 + warp] = ds;   // this chunk's table entry of my warp
        const int sslot_next = sslot + 1 == NSLOT ? 0 : sslot + 1;
        sv_pre = fetch_seam(sslot_next);
        if (LOGITS) {
            zv = zv_next;
            step0 += CH;
            zv_next = z_fetch(step0);
        }
        mbar_wait(&em_full[em_stage], (uint32_t)em_phase);
        if (n == CH) {
#pragma unroll
            for (int f = 0; f < CH; ++f) frame(f, sv, std::false_type{});
            if (is31) {
#pragma unroll
                for (int f = 0; f < CH; ++f) sv_out[sslot * SW + f] = sout[f];
            }
        } else {
#pragma unroll 1
            for (int f = 0; f < n; ++f) frame(f, sv, std::true_type{});
        }
        if (is31) {
            sv_out[sslot * SW + CH] = (float)off_me;
            sv_out[sslot * SW + CH + 1] = (float)off_row;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&em_empty[em_stage]);
        remaining -= n;
        ++chunk_no;
        if (SAVE) row_out += (int64_t)n * row_step;
        em_chunk += CH * slot_bytes;
        if (++em_stage == NST) { em_stage = 0; em_phase ^= 1; em_chunk = em_base; }
        sslot = sslot_next;
    }
    if (__any_sync(FULL, nan_acc != nan_acc) && lane == 0) p.nan_flag[b] = 1;
    // frontier row for the join kernel / the backward call (same conversion; nsteps == 0: the virtual start row)
    {
        float *fin = p.finals + ((int64_t)b * 2 + dir) * row_elems;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            fin[bpos[k]] = ab[k];
            fin[lpos[k]] = al[k];
        }
        if (lane == 0) tabs_bd[(p.NCH - 1) * 16 + warp] = ds;   // the frontier's table
        if (first && lane == 0) {
            *reinterpret_cast<double *>(fin + 2 * P_pad + 2) = (double)off_row;
            fin[2 * P_pad + 4] = (float)K;
        }
    }
}

// Join the alpha frontier (row m-1, or the virtual start row) with the beta frontier (row m):
//   log P = off_a + off_b + lse_s( lse(alpha[s], alpha[s-1], skip ? alpha[s-2]) + beta_m[s] )
__global__ void __launch_bounds__(256) ctc_join_kernel(const CtcParams p) {
    __shared__ float red_m[8], red_s[8];
    const int b = utterance_of_cta(p), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (b < 0) return;
    int L = p.tgt_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    const int P_pad = p.cfg.P_pad, row_elems = p.cfg.row_elems;
    const float *fa = p.finals + (int64_t)b * 2 * row_elems;
    const float *fb = fa + row_elems;
    const float *fal = fa + P_pad + 1, *fbl = fb + P_pad + 1;  // label p at [p]
    const int32_t *tg = p.targets + p.tgt_off[b];
    // per-warp offset tables of the wavefront forward (see "offset tables"; Kf = 0: none): every frontier is
    // taken relative to its largest table entry
    const int Kf = (int)fa[2 * P_pad + 4], span = Kf > 0 ? 32 * Kf : 1 << 30;
    const float *ta = p.tabs + (((int64_t)b * 2 + 0) * p.NCH + (p.NCH - 1)) * 16, *tb = ta + (int64_t)p.NCH * 16;
    float Da = 0.f, Db = 0.f;
    if (Kf > 0) {
        Da = ta[0];
        Db = tb[0];
        for (int w = 1; w <= L / span; ++w) { Da = fmaxf(Da, ta[w]); Db = fmaxf(Db, tb[w]); }
    }
    auto A_blank = [&](int pp) { return fa[pp] + (Kf > 0 ? ta[pp / span] - Da : 0.f); };             // alpha: element = pair
    auto A_label = [&](int li) { return fal[li] + (Kf > 0 ? ta[li / span] - Da : 0.f); };
    auto B_blank = [&](int pp) { return fb[pp] + (Kf > 0 ? tb[(L - pp) / span] - Db : 0.f); };       // beta: element = L - pair
    auto B_label = [&](int li) { return fbl[li] + (Kf > 0 ? tb[(L - li - 1) / span] - Db : 0.f); };
    float mx = kNeg, sm = 0.f;  // running max and sum of 2^(x - mx)
    auto push = [&](float x) {
        const float nm = fmaxf(mx, x);
        sm = sm * ex2_approx(mx - nm) + ex2_approx(x - nm);
        mx = nm;
    };
    for (int pp = tid; pp <= L; pp += 256) {
        const float a_b = A_blank(pp);
        const float a_lp = pp > 0 ? A_label(pp - 1) : kNeg;
        const float A = lse2(a_b, a_lp);
        push(A + B_blank(pp));
        if (pp < L) {
            bool skip = false;
            if (pp > 0) skip = tg[pp] != tg[pp - 1];
            push(lse2(A_label(pp), skip ? A : a_b) + B_label(pp));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, d), os = __shfl_xor_sync(0xffffffffu, sm, d);
        const float nm = fmaxf(mx, om);
        sm = sm * ex2_approx(mx - nm) + os * ex2_approx(om - nm);
        mx = nm;
    }
    if (lane == 0) { red_m[warp] = mx; red_s[warp] = sm; }
    // A label outside [0, V) (or equal to nothing the lattice kernels could index) is an argument error the
    // asynchronous device entry points cannot return: the kernels clamp it for memory safety and the utterance's
    // likelihood becomes NaN here (loss and gradient NaN -- loud, never a silently wrong number).
    int bad = 0;
    for (int i = tid; i < L; i += 256) bad |= (tg[i] < 0 || tg[i] >= p.V) ? 1 : 0;
    bad = __syncthreads_or(bad);
    if (tid == 0) {
        float M = red_m[0], S = red_s[0];
        for (int w = 1; w < 8; ++w) {
            const float nm = fmaxf(M, red_m[w]);
            S = S * ex2_approx(M - nm) + red_s[w] * ex2_approx(red_m[w] - nm);
            M = nm;
        }
        const bool dead = M < kNegTest;
        const double off_a = *reinterpret_cast<const double *>(fa + 2 * P_pad + 2);
        const double off_b = *reinterpret_cast<const double *>(fb + 2 * P_pad + 2);
        const double logp2 = (double)M + (double)log2f(S) + off_a + off_b + (double)Da + (double)Db;
        if ((p.abort_word && *p.abort_word != 0) || p.nan_flag[b] != 0) bad = 1;
        if (!p.join_keeps_nll)
            p.nll[b] = bad ? __int_as_float(0x7fc00000) : (dead ? __int_as_float(0x7f800000) : (float)(-logp2 * 0.6931471805599453));
        p.nll2[b] = -logp2;
    }
}

// Fused reduction ('none' | 'mean' | 'sum' | 'mean_volume') + zero_infinity + d loss / d nll_b.
__global__ void __launch_bounds__(256) ctc_reduce_kernel(const float *nll, const int32_t *tgt_len, int64_t B,
                                                         int reduction, int zero_inf, float *loss_out,
                                                         float *grad_scale) {
    __shared__ double red_a[8], red_b[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double a = 0.0, l = 0.0;
    for (int64_t b = tid; b < B; b += 256) {
        float x = nll[b];
        const bool dropped = zero_inf && x == __int_as_float(0x7f800000);   // +inf only (torch: where(loss == inf, 0, loss))
        if (dropped) x = 0.f;
        const int L = tgt_len[b];
        const float denom = (float)(L < 1 ? 1 : L);
        if (reduction == 0) loss_out[b] = x;
        float g = dropped ? 0.f : 1.f;
        if (reduction == 1) { a += (double)(x / denom); g = g / (denom * (float)B); }
        else { a += (double)x; }
        l += (double)L;
        if (grad_scale && reduction != 3) grad_scale[b] = g;
    }
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        l += __shfl_xor_sync(0xffffffffu, l, d);
    }
    if (lane == 0) { red_a[warp] = a; red_b[warp] = l; }
    __syncthreads();
    double A = 0.0, Lsum = 0.0;
    for (int w = 0; w < 8; ++w) { A += red_a[w]; Lsum += red_b[w]; }
    if (Lsum < 1.0) Lsum = 1.0;
    if (tid == 0) {
        if (reduction == 1) loss_out[0] = (float)(A / (double)B);
        else if (reduction == 2) loss_out[0] = (float)A;
        else if (reduction == 3) loss_out[0] = (float)(A / Lsum);
    }
    if (grad_scale && reduction == 3)
        for (int64_t b = tid; b < B; b += 256) {
            const bool dropped = zero_inf && nll[b] == __int_as_float(0x7f800000);
            grad_scale[b] = dropped ? 0.f : (float)(1.0 / Lsum);
        }
}

// Row normalisers for the logits entry points (the step before the path, SURVEY 8 f-1: F.log_softmax at
// modeling_wav2vec2.py:1725 / ssak/infer/general.py:99-101): zl[t,b] = -log2 sum_v exp(x[t,b,v]) for t < T_b.
// One warp per row, tiles of 256 columns, running (max, sum) merged per tile: one MUFU per element.  The log-probs
// themselves are never written: the lattice kernels form x * log2(e) + zl on the fly.
__global__ void __launch_bounds__(256) ctc_row_lse_kernel(const CtcParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= p.T * p.B) return;
    const int64_t t = row / p.B, b = row % p.B;
    int Tb = p.in_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.T ? (int)p.T : Tb);
    if (t >= Tb) {
        if (lane == 0) p.zl[row] = 0.f;
        return;
    }
    const float *x = p.lp + t * p.st + b * p.sb;
    const int V = p.V;
    float m = kNeg, s = 0.f;  // running max (log2 domain) and sum of 2^(x log2e - m)
    for (int v0 = 0; v0 < V; v0 += 256) {
        float xv[8];
        float tm = kNeg;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int v = v0 + i * 32 + lane;
            xv[i] = v < V ? fmaxf(__ldg(x + v) * kLog2e, kNeg) : kNeg;
            tm = fmaxf(tm, xv[i]);
        }
        const float nm = fmaxf(m, tm);
        float ts = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) ts += ex2_approx(xv[i] - nm);
        s = s * ex2_approx(m - nm) + ts;
        m = nm;
    }
    const float M = warp_max(m);
    s *= ex2_approx(m - M);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) p.zl[row] = -(M + lg2_approx(s));
}

// Utterance-sharded reduction: the local part of [numerator, denominator] as doubles, ready for the
// ONE all-reduce, plus d numerator / d nll_b.  reduction 1 'mean': num = sum nll_b / clamp(L_b,1), den = B_local
// (the caller divides by the a-priori known global batch); 2 'sum': num = sum nll_b, den = 1 per rank (unused);
// 3 'mean_volume': num = sum nll_b, den = sum L_b.
__global__ void __launch_bounds__(256) ctc_shard_pack_kernel(const float *nll, const int32_t *tgt_len, int64_t B,
                                                             int reduction, int zero_inf, double *packed,
                                                             float *gscale) {
    __shared__ double red_a[8], red_b[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double a = 0.0, l = 0.0;
    for (int64_t b = tid; b < B; b += 256) {
        float x = nll[b];
        const bool dropped = zero_inf && x == __int_as_float(0x7f800000);
        if (dropped) x = 0.f;
        const int L = tgt_len[b];
        const float w = reduction == 1 ? 1.0f / (float)(L < 1 ? 1 : L) : 1.0f;
        a += (double)(x * w);
        l += (double)L;
        gscale[b] = dropped ? 0.f : w;
    }
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        l += __shfl_xor_sync(0xffffffffu, l, d);
    }
    if (lane == 0) { red_a[warp] = a; red_b[warp] = l; }
    __syncthreads();
    if (tid == 0) {
        double A = 0.0, Ls = 0.0;
        for (int w = 0; w < 8; ++w) { A += red_a[w]; Ls += red_b[w]; }
        packed[0] = A;
        packed[1] = reduction == 1 ? (double)B : (reduction == 3 ? Ls : 1.0);
    }
}

// After the all-reduce: loss = num / den with den = global_batch ('mean'), 1 ('sum') or the reduced sum of
// lengths ('mean_volume'); inv_den feeds the backward scale.
__global__ void ctc_shard_finish_kernel(const double *packed, int reduction, double global_batch, float *loss,
                                        float *inv_den) {
    double den = reduction == 1 ? global_batch : (reduction == 3 ? (packed[1] < 1.0 ? 1.0 : packed[1]) : 1.0);
    loss[0] = (float)(packed[0] / den);
    inv_den[0] = (float)(1.0 / den);
}

// grad_out[b] of the backward launch = upstream gradient x d loss / d nll_b
__global__ void ctc_shard_grad_scale_kernel(const float *gscale, const float *grad_loss, const float *inv_den,
                                            int64_t B, float *out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) out[b] = gscale[b] * (grad_loss[0] * inv_den[0]);
}

// grad[t,b,:] *= per_utt[b] * s1[0] * s2[0] (absent factors = 1); a row whose factor is exactly 1 is left alone, so the
// usual case (upstream gradient 1) costs a launch and no memory traffic.  One warp per (t, b) row.
__global__ void __launch_bounds__(256) ctc_grad_scale_kernel(float *grad, int64_t T, int64_t B, int V, int64_t gst,
                                                             int64_t gsb, const float *per_utt, const float *s1,
                                                             const float *s2) {
    const int lane = threadIdx.x & 31;
    const float common = (s1 ? s1[0] : 1.f) * (s2 ? s2[0] : 1.f);
    if (!per_utt && common == 1.f) return;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < T * B; r += (int64_t)gridDim.x * 8) {
        const int64_t t = r / B, b = r - t * B;
        const float f = (per_utt ? per_utt[b] : 1.f) * common;
        if (f == 1.f) continue;
        float *g = grad + t * gst + b * gsb;
        for (int v = lane; v < V; v += 32) g[v] *= f;
    }
}

// ---------------------------------------------------------------------------- launchers
template <bool GRAD>
static int launch_lattice(const CtcParams &p, cudaStream_t stream) {
    const CtcCfg &c = p.cfg;
    const size_t smem_bytes = smem_bytes_for(c, p.V, p.Lmax, GRAD);
    if (smem_bytes > 227 * 1024) return SSAK_ERR_UNSUPPORTED;
    dim3 grid((unsigned)(p.mask ? p.n_slots : p.B), 2), block((c.W + (GRAD ? 2 + c.G + c.PW : 1)) * 32);
#define SSAK_LAUNCH4(KK, CC, ZZ, SP)                                                           \
    {                                                                                          \
        auto kern = ctc_lattice_kernel<KK, GRAD, CC, ZZ, SP>;                                  \
        cudaError_t e = ensure_max_smem<ctc_lattice_kernel<KK, GRAD, CC, ZZ, SP>>();           \
        if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }                \
        kern<<<grid, block, smem_bytes, stream>>>(p);                                          \
    }
#define SSAK_LAUNCH3(KK, CC, ZZ)                                                               \
    if (GRAD && KK <= 4 && c.PW > 0) SSAK_LAUNCH4(KK, CC, ZZ, (GRAD && KK <= 4)) else SSAK_LAUNCH4(KK, CC, ZZ, false)
#define SSAK_LAUNCH2(KK, CC)                                                                   \
    if (p.zl) { SSAK_LAUNCH3(KK, CC, true) } else { SSAK_LAUNCH3(KK, CC, false) }
#define SSAK_LAUNCH(KK)                                                                        \
    case KK:                                                                                   \
        if (c.chunk == 8) { SSAK_LAUNCH2(KK, 8) } else { SSAK_LAUNCH2(KK, 4) }                 \
        break;
    switch (c.K) {
        SSAK_LAUNCH(1)
        SSAK_LAUNCH(2)
        SSAK_LAUNCH(4)
        SSAK_LAUNCH(8)
        default: return SSAK_ERR_UNSUPPORTED;
    }
#undef SSAK_LAUNCH4
#undef SSAK_LAUNCH3
#undef SSAK_LAUNCH2
#undef SSAK_LAUNCH
    return check_launch();
}

// The wavefront forward kernel when the shape fits it (SSAK_ERR_UNSUPPORTED -> the barrier kernel is used).
static int launch_forward_wave(const CtcParams &p, cudaStream_t stream) {
    // Used on at most one CTA per SM (2B <= 148: latency-bound, the wavefront is ~25 % faster); with several CTAs per SM
    // the barrier kernel's fatter lanes issue fewer instructions per cell (1.25 vs 1.54 ms at B = 1024).  Rows
    // saved for the backward come with per-warp offset tables, which only the posterior warps (cfg.PW > 0) read.
    // SSAK_CTC_FWD_WAVE=1 forces it where it is valid, =0 disables it.
    const int mode = env_int("SSAK_CTC_FWD_WAVE", -1);
    const int sms = device_sm_count();
    if (mode == 0 || (mode < 0 && 2 * p.B > sms)) return SSAK_ERR_UNSUPPORTED;   // B = 128: 0.243 vs 0.232 ms
    if (p.rows != nullptr && p.cfg.PW == 0) return SSAK_ERR_UNSUPPORTED;
    const int64_t P = (int64_t)p.Lmax + 1;
    if (P > 1024 || p.T > 100000) return SSAK_ERR_UNSUPPORTED;   // 32 * 4 * 8 chain elements; float-exact offsets
    FwdWaveCfg wc;
    // A lone warp issues ~0.3 instructions per cycle whatever its ILP (measured), so latency-bound shapes want MANY
    // thin warps: one chain element per lane up to 16 warps, then 2, then 4 per lane.
    wc.K = env_int("SSAK_CTC_FWD_K", P <= 512 ? 1 : 4);
    if (wc.K != 1 && wc.K != 2 && wc.K != 4) return SSAK_ERR_UNSUPPORTED;
    wc.W = (int)((P + 32 * wc.K - 1) / (32 * wc.K));
    if (wc.W > (wc.K == 1 ? 16 : 8)) return SSAK_ERR_UNSUPPORTED;
    wc.slot_bytes = ring_slot_bytes(p.V);
    wc.chunk = 8 * wc.slot_bytes <= 16384 ? 8 : 4;
    const int budget = p.B <= sms ? 100 * 1024 : 40 * 1024;
    int stages = budget / (wc.chunk * wc.slot_bytes);
    if (stages > wc.W + 6) stages = wc.W + 6;
    stages = env_int("SSAK_CTC_FWD_STAGES", stages);
    if (stages < wc.W + 2 || stages > 24) return SSAK_ERR_UNSUPPORTED;
    wc.stages = stages;
    const size_t smem_bytes = (size_t)fwd_wave_smem(wc.W, wc.stages, wc.chunk, wc.slot_bytes).total;
    if (smem_bytes > 227 * 1024) return SSAK_ERR_UNSUPPORTED;
    dim3 grid((unsigned)(p.mask ? p.n_slots : p.B), 2), block((wc.W + 1) * 32);
#define SSAK_FW4(KK, CC, ZZ, SS)                                                               \
    {                                                                                          \
        auto kern = ctc_forward_wave_kernel<KK, CC, ZZ, SS>;                                   \
        cudaError_t e = ensure_max_smem<ctc_forward_wave_kernel<KK, CC, ZZ, SS>>();            \
        if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }                \
        kern<<<grid, block, smem_bytes, stream>>>(p, wc);                                      \
    }
#define SSAK_FW3(KK, CC, ZZ) if (p.rows) SSAK_FW4(KK, CC, ZZ, true) else SSAK_FW4(KK, CC, ZZ, false)
#define SSAK_FW2(KK, CC) if (p.zl) { SSAK_FW3(KK, CC, true) } else { SSAK_FW3(KK, CC, false) }
#define SSAK_FW(KK) if (wc.chunk == 8) { SSAK_FW2(KK, 8) } else { SSAK_FW2(KK, 4) }
    if (wc.K == 1) { SSAK_FW(1) } else if (wc.K == 2) { SSAK_FW(2) } else { SSAK_FW(4) }
#undef SSAK_FW
#undef SSAK_FW2
#undef SSAK_FW3
#undef SSAK_FW4
    return check_launch();
}

// Throughput kernels (ctc_lin32.cu: one warp per (utterance, direction), linear domain, block floating point, no
// stored lattice): used when they cover the shape (V <= 128, targets up to 415 labels) AND the batch fills the GPU --
// a chain is one warp, so below ~3 chains per SM the latency-tuned log-domain kernels (several warps per utterance)
// are faster (B = 64: 0.42 vs 1.1 ms; B = 256 ragged: 1.15 vs 1.06 ms of kernels, and 1.5e-4 vs 7e-7 of gradient
// error on unpeaked emissions; B = 1024: 4.1 vs 2.4 ms).
// SSAK_CTC_LIN32=1 forces them wherever they are valid, =0 disables them.
// `aligned`: rows of log-probabilities 16-byte aligned and V % 4 == 0 (what the large-vocabulary kernels need for their
// bulk copies); -1: not known (workspace queries).
// `saved` == false (a forward-only call): never -- their likelihood is only verified by backward()'s self-check
// (states lost to the fp32 range show up as posterior mass that does not sum to 1), so a forward() nobody follows
// up must come from the log-domain kernels, whose range is unlimited.
static int lin_k(int64_t Lmax, int64_t V, int64_t B, int aligned, bool saved) {
    const int mode = env_int("SSAK_CTC_LIN32", -1);
    if (mode == 0 || !saved) return 0;
    if (mode < 0 && 2 * B < 3 * (int64_t)device_sm_count()) return 0;   // B >= 1.5 x SMs (222 on a B200)
    if (V > lin32::MAXV) {
        // large vocabularies (ctc_lin32_lv.cuh): whole rows in a per-warp ring leave room for 7 chains per SM, so the
        // gather kernels only pay off with many chains (B = 512, V = 1024 [C5]: 1.27 vs 1.22 ms; B = 1024: 2.25 vs
        // 2.19 ms; B = 2048, T = 400, V = 512: 1.25 vs 2.11 ms)
        if (aligned == 0 || (V & 3) != 0) return 0;
        if (mode < 0 && B < 6 * (int64_t)device_sm_count()) return 0;
    }
    return lin32::lanes_k(Lmax, V);
}
static inline int rows_aligned(const float *lp, int64_t st, int64_t sb, int64_t V) {
    return ((reinterpret_cast<uintptr_t>(lp) & 15) == 0 && ((st | sb | V) & 3) == 0) ? 1 : 0;
}
// row blocks kept for utterances the throughput kernels hand back to the log-domain ones (fp32 range, see
// ctc_lin32.cu): every utterance of a small batch, 1/8 of a large one (beyond: NaN likelihood / gradient, loud)
static inline int64_t lin_slots(int64_t B) { return B <= 32 ? B : std::max<int64_t>(32, B / 8); }

struct WsLayout { size_t nll2, abort_word, finals, zl, tabs, lin_fr, lin_ck, lin_order, rows, total; };
static inline int tab_slots(int64_t T) { return (int)((T + 7) / 8) + 2; }  // >= chunks of T/2 frames (chunk >= 4) + 1
// V < 0: the vocabulary is not known (ssak_ctc_loss_workspace_bytes); aligned < 0 with V > 128: the alignment of the
// rows is not known (ssak_ctc_loss_workspace_bytes_v) -- in both cases room for either kernel family
static WsLayout ws_layout(int64_t T, int64_t B, int64_t V, int64_t Lmax, int row_elems, bool saved, int aligned) {
    WsLayout w;
    size_t o = 0;
    const int K = lin_k(Lmax, V < 0 ? 1 : V, B, aligned, saved);
    const bool either = K > 0 && (V < 0 || (V > lin32::MAXV && aligned < 0));
    w.nll2 = o;   o += align_up((size_t)B * sizeof(double), 256);
    // abort word (+ the slot counter of the throughput mode at +4), then nan_flag[B], then the throughput kernels'
    // flags[B], slot[B] and slot_b[n_slots <= B]: one memset
    w.abort_word = o; o += 256 + 4 * align_up((size_t)B * sizeof(int), 256);
    w.lin_fr = w.lin_ck = w.lin_order = o;
    if (K > 0) {
        const size_t ck_row = (size_t)lin32::ck_row_elems(K);
        w.lin_fr = o; o += align_up((size_t)B * 2 * ck_row * sizeof(float), 256);
        w.lin_ck = o;
        if (saved) o += align_up((size_t)B * 2 * lin32::n_checkpoints(T) * ck_row * sizeof(float), 256);
        w.lin_order = o; o += align_up((size_t)B * sizeof(int), 256);
    }
    w.finals = o; o += align_up((size_t)B * 2 * row_elems * sizeof(float), 256);
    w.zl = o;     o += align_up((size_t)B * (size_t)T * sizeof(float), 256);   // row normalisers (logits entry points)
    w.tabs = o;   o += align_up((size_t)B * 2 * tab_slots(T) * 16 * sizeof(float), 256);
    // log-domain half lattices: one row block per utterance -- in the throughput mode only for the few utterances
    // that may be handed back (lin_slots)
    const int64_t blocks = (K > 0 && !either) ? lin_slots(B) : B;
    w.rows = o;   if (saved) o += align_up((size_t)blocks * (size_t)T * row_elems * sizeof(float), 256);
    w.total = o + 256;
    return w;
}

static int fill_params(CtcParams *p, const float *log_probs, int64_t T, int64_t B, int64_t V,
                       int64_t st, int64_t sb, const int32_t *targets, const int64_t *tgt_off,
                       const int32_t *in_len, const int32_t *tgt_len, int64_t Lmax, int32_t blank,
                       void *workspace, size_t workspace_bytes, bool saved, bool logits) {
    if (!log_probs || !targets || !tgt_off || !in_len || !tgt_len || !workspace)
        return SSAK_ERR_INVALID_ARGUMENT;
    if (T < 0 || B <= 0 || V <= 0 || Lmax < 0 || blank < 0 || blank >= V || T > 0x7ffffff0 ||
        V > (1 << 20) || B > 65535 * 32)
        return SSAK_ERR_INVALID_ARGUMENT;
    if (T > 300000) return SSAK_ERR_UNSUPPORTED;  // re-centring offsets are kept exact as fp32 integers (< 2^24)
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return SSAK_ERR_INVALID_ARGUMENT;
    if (!choose_cfg(Lmax, B, (int)V, &p->cfg)) return SSAK_ERR_UNSUPPORTED;
    const WsLayout w = ws_layout(T, B, V, Lmax, p->cfg.row_elems, saved, rows_aligned(log_probs, st, sb, V));
    if (workspace_bytes < w.total) return SSAK_ERR_WORKSPACE;
    p->lp = log_probs; p->T = T; p->B = B; p->V = (int)V; p->st = st; p->sb = sb;
    p->targets = targets; p->tgt_off = tgt_off; p->in_len = in_len; p->tgt_len = tgt_len;
    p->Lmax = (int)Lmax; p->blank = blank;
    char *ws = reinterpret_cast<char *>(workspace);
    p->nll2 = reinterpret_cast<double *>(ws + w.nll2);
    p->finals = reinterpret_cast<float *>(ws + w.finals);
    p->rows = saved ? reinterpret_cast<float *>(ws + w.rows) : nullptr;
    p->zl = logits ? reinterpret_cast<float *>(ws + w.zl) : nullptr;
    p->tabs = reinterpret_cast<float *>(ws + w.tabs);
    p->abort_word = reinterpret_cast<int *>(ws + w.abort_word);
    p->nan_flag = reinterpret_cast<int *>(ws + w.abort_word + 256);
    p->mask = nullptr;
    p->mask_bits = 1;
    p->slot_b = nullptr; p->slot_count = nullptr; p->n_slots = 0;
    p->join_keeps_nll = 0;
    p->slot = nullptr;
    p->NCH = tab_slots(T);
    p->nll = nullptr; p->grad_out = nullptr; p->grad = nullptr; p->gst = p->gsb = 0;
    p->zero_inf = 0;
    return SSAK_OK;
}

}  // namespace ssak

using namespace ssak;

extern "C" size_t ssak_ctc_loss_workspace_bytes(int64_t T, int64_t B, int64_t max_target_len,
                                                int save_for_backward) {
    CtcCfg c;
    if (T < 0 || T > 300000 || B <= 0 || max_target_len < 0 || !choose_cfg(max_target_len, B, 64, &c)) return 0;
    return ws_layout(T, B, -1, max_target_len, c.row_elems, save_for_backward != 0, -1).total;
}

/* The same with the vocabulary size known: the throughput kernels (V <= 128, targets up to 415 labels) keep
 * checkpoints instead of half lattices, ~8x less workspace. */
extern "C" size_t ssak_ctc_loss_workspace_bytes_v(int64_t T, int64_t B, int64_t V, int64_t max_target_len,
                                                  int save_for_backward) {
    CtcCfg c;
    if (T < 0 || T > 300000 || B <= 0 || V <= 0 || max_target_len < 0 || !choose_cfg(max_target_len, B, (int)V, &c))
        return 0;
    return ws_layout(T, B, V, max_target_len, c.row_elems, save_for_backward != 0, -1).total;
}

// Fill the parameters of the throughput kernels from the log-domain ones (same problem, same workspace).
static bool lin_params(const CtcParams &p, void *workspace, bool saved, lin32::Params *q) {
    const int al = rows_aligned(p.lp, p.st, p.sb, p.V);
    const int K = lin_k(p.Lmax, p.V, p.B, al, saved);
    if (K == 0) return false;
    const WsLayout w = ws_layout(p.T, p.B, p.V, p.Lmax, p.cfg.row_elems, saved, al);
    char *ws = reinterpret_cast<char *>(workspace);
    q->lp = p.lp; q->T = p.T; q->B = p.B; q->V = p.V; q->st = p.st; q->sb = p.sb;
    q->targets = p.targets; q->tgt_off = p.tgt_off; q->in_len = p.in_len; q->tgt_len = p.tgt_len;
    q->Lmax = p.Lmax; q->blank = p.blank; q->zl = p.zl;
    q->K = K;
    q->ck_row = lin32::ck_row_elems(K);
    q->NCK = lin32::n_checkpoints(p.T);
    q->fr = reinterpret_cast<float *>(ws + w.lin_fr);
    q->ck = reinterpret_cast<float *>(ws + w.lin_ck);
    q->nll2 = p.nll2; q->nll = p.nll;
    q->flags = reinterpret_cast<int *>(ws + w.abort_word + 256 + align_up((size_t)p.B * sizeof(int), 256));
    q->slot = reinterpret_cast<int *>(ws + w.abort_word + 256 + 2 * align_up((size_t)p.B * sizeof(int), 256));
    q->slot_b = reinterpret_cast<int *>(ws + w.abort_word + 256 + 3 * align_up((size_t)p.B * sizeof(int), 256));
    q->slot_counter = reinterpret_cast<int *>(ws + w.abort_word + 4);
    q->order = lin32::ordered(p.B) ? reinterpret_cast<int *>(ws + w.lin_order) : nullptr;
    q->n_slots = saved ? (int)lin_slots(p.B) : (int)p.B;   // (no rows are stored without save_for_backward)
    q->grad_out = p.grad_out; q->grad = p.grad; q->gst = p.gst; q->gsb = p.gsb; q->zero_inf = p.zero_inf;
    q->save = saved ? 1 : 0;
    q->mass_tol = 1e-6f * (float)env_int("SSAK_LIN32_TOL_PPM", 20);
    return true;
}

/* 1 when the kernels cover the shape, 0 otherwise (max_target_len > 4095, T > 300000, B <= 0, or a vocabulary
 * whose rows do not fit the shared-memory emission ring, V > ~2040): callers that replace a generic operator use it
 * to delegate unsupported shapes instead of failing. */
extern "C" int ssak_ctc_loss_supported(int64_t T, int64_t B, int64_t V, int64_t max_target_len) {
    CtcCfg c;
    if (T < 0 || T > 300000 || B <= 0 || V <= 0 || V > (1 << 20) || max_target_len < 0) return 0;
    if (!choose_cfg(max_target_len, B, (int)V, &c)) return 0;
    return smem_bytes_for(c, (int)V, (int)max_target_len, true) <= (size_t)kMaxDynSmem ? 1 : 0;
}

/* Diagnostics: which kernel family computed each utterance in the last forward (+ backward) call on this workspace.
 * flags_out[b] (device, int32): 0 throughput kernels; bit 0: handed to the log-domain kernels by forward(); bit 1: by
 * backward() (its self-check failed); bit 2: nobody (no row block left: NaN).  All zeros when the shape is not
 * covered by the throughput kernels. */
extern "C" int ssak_ctc_loss_path_flags(const void *workspace, int64_t T, int64_t B, int64_t V, int64_t max_target_len,
                                        int32_t save_for_backward, int32_t *flags_out, ssak_stream_t stream) {
    CtcCfg c;
    if (!workspace || !flags_out || B <= 0 || V <= 0 || !choose_cfg(max_target_len, B, (int)V, &c))
        return SSAK_ERR_INVALID_ARGUMENT;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e;
    {   // (the flag words sit at the same offset whichever kernels ran; forward() clears them)
        const WsLayout w = ws_layout(T, B, V, max_target_len, c.row_elems, save_for_backward != 0, -1);
        const char *ws = reinterpret_cast<const char *>(workspace);
        e = cudaMemcpyAsync(flags_out, ws + w.abort_word + 256 + align_up((size_t)B * sizeof(int), 256),
                            (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToDevice, s);
    }
    if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }
    return SSAK_OK;
}

// kernels of ours per forward + backward call pair (what bench.py reports as gpu_launches)
extern "C" int ssak_ctc_loss_launches(int64_t B, int64_t V, int64_t max_target_len, int32_t logits) {
    // throughput mode: memset excluded; forward: [row lse] [order] lin32 fwd + lin32 join + masked log-domain fwd +
    // join; backward: lin32 bwd + masked log-domain fwd + join + bwd + orphans.  Log-domain mode: fwd + join, bwd.
    return (logits ? 1 : 0) + (lin_k(max_target_len, V, B, 1, true) > 0 ? 9 + (lin32::ordered(B) ? 1 : 0) : 3);
}

static int forward_impl(const float *x, int64_t T, int64_t B, int64_t V, int64_t st, int64_t sb,
                        const int32_t *targets, const int64_t *target_offsets, const int32_t *input_lengths,
                        const int32_t *target_lengths, int64_t max_target_len, int32_t blank,
                        int32_t save_for_backward, float *neg_log_likelihood, void *workspace,
                        size_t workspace_bytes, ssak_stream_t stream, bool logits) {
    CtcParams p;
    if (!neg_log_likelihood) return SSAK_ERR_INVALID_ARGUMENT;
    int rc = fill_params(&p, x, T, B, V, st, sb, targets, target_offsets, input_lengths, target_lengths,
                         max_target_len, blank, workspace, workspace_bytes, save_for_backward != 0, logits);
    if (rc != SSAK_OK) return rc;
    p.nll = neg_log_likelihood;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (logits && T > 0) {
        ctc_row_lse_kernel<<<(unsigned)((T * B + 7) / 8), 256, 0, s>>>(p);
        rc = check_launch();
        if (rc != SSAK_OK) return rc;
    }
    {
        cudaError_t e = cudaMemsetAsync(p.abort_word, 0, 256 + 4 * align_up((size_t)B * sizeof(int), 256), s);
        if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }
    }
    // Throughput kernels first (no stored lattice); the utterances they hand back (flags, see ctc_lin32.cu) are
    // then recomputed by the log-domain kernels below, launched over the same grid with a mask.
    lin32::Params q;
    if (lin_params(p, workspace, save_for_backward != 0, &q)) {
        rc = lin32::launch_forward(q, s);
        if (rc != SSAK_OK) return rc;
        p.mask = q.flags;
        p.mask_bits = 1;
        p.slot = q.slot;
        p.slot_b = q.slot_b; p.slot_count = q.slot_counter; p.n_slots = q.n_slots;
    }
    rc = launch_forward_wave(p, s);
    if (rc == SSAK_ERR_UNSUPPORTED) rc = launch_lattice<false>(p, s);
    if (rc != SSAK_OK) return rc;
    ctc_join_kernel<<<(unsigned)(p.mask ? p.n_slots : B), 256, 0, s>>>(p);
    return check_launch();
}

static int backward_impl(const float *grad_out, const float *x, int64_t T, int64_t B, int64_t V, int64_t st,
                         int64_t sb, const int32_t *targets, const int64_t *target_offsets,
                         const int32_t *input_lengths, const int32_t *target_lengths, int64_t max_target_len,
                         int32_t blank, int32_t zero_infinity, const float *neg_log_likelihood, float *grad,
                         int64_t g_stride_t, int64_t g_stride_b, void *workspace, size_t workspace_bytes,
                         ssak_stream_t stream, bool logits) {
    CtcParams p;
    if (!grad_out || !neg_log_likelihood || !grad) return SSAK_ERR_INVALID_ARGUMENT;
    int rc = fill_params(&p, x, T, B, V, st, sb, targets, target_offsets, input_lengths, target_lengths,
                         max_target_len, blank, workspace, workspace_bytes, true, logits);
    if (rc != SSAK_OK) return rc;
    p.nll = const_cast<float *>(neg_log_likelihood);
    p.grad_out = grad_out; p.grad = grad; p.gst = g_stride_t; p.gsb = g_stride_b;
    p.zero_inf = zero_infinity;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    lin32::Params q;
    if (lin_params(p, workspace, true, &q)) {
        rc = lin32::launch_backward(q, s);
        if (rc != SSAK_OK) return rc;
        // utterances whose self-check failed in this call (flag bit 1): the log-domain forward first (their half
        // lattices were never stored), then the log-domain backward for them and for the ones forward() handed back
        p.mask = q.flags;
        p.slot = q.slot;
        p.slot_b = q.slot_b; p.slot_count = q.slot_counter; p.n_slots = q.n_slots;
        p.mask_bits = 2;
        // (the join rewrites nll[b] of these utterances: the throughput kernels' likelihood was provisional, and the
        //  self-check that just failed says it may have lost states to the fp32 range)
        p.join_keeps_nll = 0;
        rc = launch_forward_wave(p, s);
        if (rc == SSAK_ERR_UNSUPPORTED) rc = launch_lattice<false>(p, s);
        if (rc != SSAK_OK) return rc;
        ctc_join_kernel<<<(unsigned)(p.mask ? p.n_slots : B), 256, 0, s>>>(p);
        rc = check_launch();
        if (rc != SSAK_OK) return rc;
        p.mask_bits = 3;
        rc = launch_lattice<true>(p, s);
        if (rc != SSAK_OK) return rc;
        return lin32::launch_orphans(q, s);
    }
    return launch_lattice<true>(p, s);
}

extern "C" int ssak_ctc_loss_forward(const float *log_probs, int64_t T, int64_t B, int64_t V,
                                     int64_t lp_stride_t, int64_t lp_stride_b,
                                     const int32_t *targets, const int64_t *target_offsets,
                                     const int32_t *input_lengths, const int32_t *target_lengths,
                                     int64_t max_target_len, int32_t blank,
                                     int32_t save_for_backward, float *neg_log_likelihood,
                                     void *workspace, size_t workspace_bytes,
                                     ssak_stream_t stream) {
    return forward_impl(log_probs, T, B, V, lp_stride_t, lp_stride_b, targets, target_offsets, input_lengths,
                        target_lengths, max_target_len, blank, save_for_backward, neg_log_likelihood, workspace,
                        workspace_bytes, stream, false);
}

extern "C" int ssak_ctc_loss_backward(const float *grad_out, const float *log_probs, int64_t T,
                                      int64_t B, int64_t V, int64_t lp_stride_t,
                                      int64_t lp_stride_b, const int32_t *targets,
                                      const int64_t *target_offsets, const int32_t *input_lengths,
                                      const int32_t *target_lengths, int64_t max_target_len,
                                      int32_t blank, int32_t zero_infinity,
                                      const float *neg_log_likelihood, float *grad,
                                      int64_t g_stride_t, int64_t g_stride_b, void *workspace,
                                      size_t workspace_bytes, ssak_stream_t stream) {
    return backward_impl(grad_out, log_probs, T, B, V, lp_stride_t, lp_stride_b, targets, target_offsets,
                         input_lengths, target_lengths, max_target_len, blank, zero_infinity, neg_log_likelihood,
                         grad, g_stride_t, g_stride_b, workspace, workspace_bytes, stream, false);
}

extern "C" int ssak_ctc_logits_forward(const float *logits, int64_t T, int64_t B, int64_t V,
                                       int64_t stride_t, int64_t stride_b, const int32_t *targets,
                                       const int64_t *target_offsets, const int32_t *input_lengths,
                                       const int32_t *target_lengths, int64_t max_target_len, int32_t blank,
                                       int32_t save_for_backward, float *neg_log_likelihood, void *workspace,
                                       size_t workspace_bytes, ssak_stream_t stream) {
    return forward_impl(logits, T, B, V, stride_t, stride_b, targets, target_offsets, input_lengths,
                        target_lengths, max_target_len, blank, save_for_backward, neg_log_likelihood, workspace,
                        workspace_bytes, stream, true);
}

extern "C" int ssak_ctc_logits_backward(const float *grad_out, const float *logits, int64_t T, int64_t B,
                                        int64_t V, int64_t stride_t, int64_t stride_b, const int32_t *targets,
                                        const int64_t *target_offsets, const int32_t *input_lengths,
                                        const int32_t *target_lengths, int64_t max_target_len, int32_t blank,
                                        int32_t zero_infinity, const float *neg_log_likelihood, float *grad,
                                        int64_t g_stride_t, int64_t g_stride_b, void *workspace,
                                        size_t workspace_bytes, ssak_stream_t stream) {
    return backward_impl(grad_out, logits, T, B, V, stride_t, stride_b, targets, target_offsets, input_lengths,
                         target_lengths, max_target_len, blank, zero_infinity, neg_log_likelihood, grad,
                         g_stride_t, g_stride_b, workspace, workspace_bytes, stream, true);
}

extern "C" int ssak_ctc_loss_reduce(const float *neg_log_likelihood, const int32_t *target_lengths,
                                    int64_t B, int32_t reduction, int32_t zero_infinity, float *loss_out,
                                    float *grad_scale, ssak_stream_t stream) {
    if (!neg_log_likelihood || !target_lengths || !loss_out || B <= 0 || reduction < 0 || reduction > 3)
        return SSAK_ERR_INVALID_ARGUMENT;
    ctc_reduce_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        neg_log_likelihood, target_lengths, B, reduction, zero_infinity, loss_out, grad_scale);
    return check_launch();
}

extern "C" int ssak_ctc_shard_pack(const float *neg_log_likelihood, const int32_t *target_lengths, int64_t B,
                                   int32_t reduction, int32_t zero_infinity, double *packed, float *grad_scale,
                                   ssak_stream_t stream) {
    if (!neg_log_likelihood || !target_lengths || !packed || !grad_scale || B <= 0 || reduction < 1 || reduction > 3)
        return SSAK_ERR_INVALID_ARGUMENT;
    ctc_shard_pack_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        neg_log_likelihood, target_lengths, B, reduction, zero_infinity, packed, grad_scale);
    return check_launch();
}

extern "C" int ssak_ctc_shard_finish(const double *packed, int32_t reduction, int64_t global_batch, float *loss_out,
                                     float *inv_den_out, ssak_stream_t stream) {
    if (!packed || !loss_out || !inv_den_out || reduction < 1 || reduction > 3 || global_batch <= 0)
        return SSAK_ERR_INVALID_ARGUMENT;
    ctc_shard_finish_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(packed, reduction,
                                                                                 (double)global_batch, loss_out,
                                                                                 inv_den_out);
    return check_launch();
}

extern "C" int ssak_ctc_shard_grad_scale(const float *grad_scale, const float *grad_loss, const float *inv_den,
                                         int64_t B, float *grad_out, ssak_stream_t stream) {
    if (!grad_scale || !grad_loss || !inv_den || !grad_out || B <= 0) return SSAK_ERR_INVALID_ARGUMENT;
    ctc_shard_grad_scale_kernel<<<(unsigned)((B + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        grad_scale, grad_loss, inv_den, B, grad_out);
    return check_launch();
}

extern "C" int ssak_ctc_loss_nll_is_provisional(int64_t B, int64_t V, int64_t max_target_len) {
    if (B <= 0 || V <= 0 || max_target_len < 0) return 0;
    return lin_k(max_target_len, V, B, 1, true) > 0 ? 1 : 0;
}

extern "C" int ssak_ctc_grad_scale(float *grad, int64_t T, int64_t B, int64_t V, int64_t g_stride_t, int64_t g_stride_b,
                                   const float *per_utterance, const float *scalar_a, const float *scalar_b,
                                   ssak_stream_t stream) {
    if (!grad || T < 0 || B <= 0 || V <= 0) return SSAK_ERR_INVALID_ARGUMENT;
    if (T == 0) return SSAK_OK;
    const int64_t rows = T * B;
    const unsigned grid = (unsigned)std::min<int64_t>((rows + 7) / 8, 64 * (int64_t)device_sm_count());
    ctc_grad_scale_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        grad, T, B, (int)V, g_stride_t, g_stride_b, per_utterance, scalar_a, scalar_b);
    return check_launch();
}
