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
//     access is a coalesced 128-byte line, and only one value per warp crosses warps per
//     frame (shared memory, double buffered, one CTA barrier per frame).
//   * Emission rows of the next frames are prefetched into a shared-memory ring with
//     cp.async.bulk (1-D TMA, completion on an mbarrier) by a dedicated producer warp and
//     gathered at the label columns.
//   * Numerics: every 8 frames the row is re-centred on its maximum (one integer REDUX per
//     warp) and the subtracted amount is accumulated in fp64, so the fp32 state values stay
//     O(10..100) instead of O(T): the rounding noise of the recursion drops by ~100x
//     compared with a plain fp32 log-domain recursion (what ATen does).
//   * Backward: dedicated gradient warps run one frame behind the recursion warps.  The
//     recursion warps publish per-state posteriors (shared memory), blank posteriors are
//     summed with an integer warp reduction in 2^-30 fixed point (deterministic), label
//     posteriors are summed by the gradient warps through a per-utterance CSR
//     (label -> positions) built once, and full gradient rows are written coalesced.
#include "common.cuh"

namespace ssak {

constexpr int kRecenter = 8;  // frames between two re-centrings of the lattice row

struct CtcCfg {
    int K;       // pairs per lane
    int W;       // recursion warps
    int G;       // gradient warps (backward kernel only)
    int P_pad;   // pair capacity = 32*K*W
    int chunk;   // frames per ring stage
    int stages;  // ring stages
    int slot_bytes;
};

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
    double *nll2;     // [B]    -log2 P (fp64: sum of the offsets + joined frontier)
    double *off_fin;  // [B][2] accumulated re-centring offset of each frontier row
    float *finals;    // [B][2][2*P_pad] frontier rows (natural state order: blanks, then labels)
    double *off_rows; // [B][T] offset of every stored row            (saved for backward)
    float *rows;      // [B][T][2*P_pad] half lattices                (saved for backward)
    float *nll;       // [B] out / in
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
static bool choose_cfg(int64_t Lmax, int64_t B, int V, CtcCfg *c) {
    const int64_t P = Lmax + 1;
    // Few CTAs (latency regime): more warps, fewer pairs per lane.  Many CTAs (throughput
    // regime): fat lanes, few warps, so several utterances share an SM.
    int wtarget = (2 * B <= 2 * 148) ? 8 : ((2 * B <= 6 * 148) ? 4 : 2);
    wtarget = env_int("SSAK_CTC_WARPS", wtarget);
    int K = env_int("SSAK_CTC_K", 0);
    if (K == 0) {
        K = 1;
        while (K < 8 && (P + 32 * K - 1) / (32 * K) > wtarget) K *= 2;
    }
    if (K != 1 && K != 2 && K != 4 && K != 8) return false;
    const int64_t W = (P + 32 * K - 1) / (32 * K);
    if (W > 16) return false;  // L <= 4095
    c->K = K;
    c->W = (int)W;
    c->P_pad = 32 * K * (int)W;
    c->G = V <= 64 ? 2 : 4;
    c->slot_bytes = ring_slot_bytes(V);
    int chunk = 8;
    while (chunk > 1 && chunk * c->slot_bytes > 16384) chunk >>= 1;
    c->chunk = chunk;
    c->stages = 4;
    return true;
}

// shared memory: [mbarriers 64][xchg 2x16 f32][wmax 16 f32][blank_acc 2 u32 (+pad)] ring | wlab ...
constexpr int kSmemXchg = 64, kSmemWmax = 64 + 128, kSmemBlank = 64 + 128 + 64, kSmemRing = 272;
static size_t smem_bytes_for(const CtcCfg &c, int V, int Lmax, bool grad) {
    size_t o = kSmemRing + (size_t)c.stages * c.chunk * c.slot_bytes;
    if (grad)
        o += 2 * (size_t)c.P_pad * sizeof(float) + ((size_t)V + 1) * sizeof(int) + (size_t)V * sizeof(int) +
             (size_t)(Lmax > 0 ? Lmax : 1) * sizeof(int);
    return align_up(o, 16);
}

// ------------------------------------------------------------------------------ kernel
// Warp roles: [0, W) recursion, W producer (bulk copies), (W, W+G] gradient (backward only).
template <int K, bool GRAD>
__global__ void __launch_bounds__(GRAD ? 672 : 544, 1) ctc_lattice_kernel(const CtcParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const CtcCfg &c = p.cfg;
    const int b = blockIdx.x;
    const int dir = blockIdx.y;  // 0: alpha (forward in time), 1: beta (backward in time)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = c.W;
    const bool compute = warp < W;
    const bool producer = warp == W;
    const unsigned FULL = 0xffffffffu;

    int Tb = p.in_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.T ? (int)p.T : Tb);
    int L = p.tgt_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    const int m = Tb >> 1;
    const int n1 = dir ? Tb - m : m;        // frames this direction owns in forward()
    const int tau0 = GRAD ? n1 : 0;         // first direction-local step of this launch
    const int nsteps = GRAD ? Tb - n1 : n1;
    const int P_pad = c.P_pad;
    const int V = p.V;
    const float *lp_b = p.lp + (int64_t)b * p.sb;
    const int32_t *tg = p.targets + p.tgt_off[b];
    const int t_first = dir ? Tb - 1 - tau0 : tau0;   // frame of step 0
    const int dt = dir ? -1 : 1;

    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    float *xchg = reinterpret_cast<float *>(smem + kSmemXchg);            // [2][16]
    float *wmax = reinterpret_cast<float *>(smem + kSmemWmax);            // [16]
    unsigned *blank_acc = reinterpret_cast<unsigned *>(smem + kSmemBlank);  // [2]
    RowRing ring;
    ring.slots = smem + kSmemRing;
    ring.full = full;
    ring.chunk = c.chunk;
    ring.stages = c.stages;
    ring.slot_bytes = c.slot_bytes;
    ring.row_bytes = 4 * V;
    float *wlab = reinterpret_cast<float *>(smem + kSmemRing + (size_t)c.stages * c.chunk * c.slot_bytes);
    int *occ_start = reinterpret_cast<int *>(wlab + 2 * P_pad);
    int *cursor = occ_start + (V + 1);
    int *occ_pos = cursor + V;

    // ---- gradient prologue: trivial outcomes ----
    double nll2 = 0.0;
    float gs = 0.f;
    if (GRAD) {
        const float nll = p.nll[b];
        gs = p.grad_out[b];
        const bool infeasible = !(nll < 3.0e38f);  // +inf (or NaN)
        if (infeasible || Tb == 0) {
            // zero_infinity: every row 0.  Otherwise torch yields NaN for t < T_b.
            if (dir == 0) {
                const float fillv = (infeasible && !p.zero_inf) ? __int_as_float(0x7fc00000) : 0.f;
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

    if (tid == 0) {
        for (int s = 0; s < c.stages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        blank_acc[0] = 0u;
        blank_acc[1] = 0u;
    }

    // ---- per-thread static data: labels of my K pairs (direction-local order) ----
    int lab[K];
    unsigned skipmask = 0;
    const int pbase = warp * 32 * K + lane;
    if (compute) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int pp = pbase + k * 32;  // direction-local pair index
            int l = p.blank;
            if (pp < L) {
                const int li = dir ? L - 1 - pp : pp;  // natural label index
                l = tg[li];
                l = l < 0 ? 0 : (l >= V ? V - 1 : l);
                if (pp >= 1) {
                    int lprev = tg[dir ? li + 1 : li - 1];
                    lprev = lprev < 0 ? 0 : (lprev >= V ? V - 1 : lprev);
                    if (lprev != l) skipmask |= 1u << k;
                }
            }
            lab[k] = l;
        }
    }

    // ---- CSR label -> natural positions (backward only; deterministic order) ----
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
                    occ_pos[base + rank] = i;
                    __syncwarp(act);
                    if (rank == 0) cursor[l] = base + __popc(mm);
                }
                __syncwarp();
            }
        }
    }

    // ---- recursion state: virtual start row (forward) or the stored frontier (backward) ----
    float ab[K], al[K];
    double off_mine = 0.0;  // accumulated re-centring offset: true value = state + off_mine
    if (compute) {
        const float *fin = p.finals + ((int64_t)b * 2 + dir) * 2 * P_pad;
        if (GRAD) off_mine = p.off_fin[(int64_t)b * 2 + dir];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int pp = pbase + k * 32;
            if (GRAD) {
                const int pb = dir ? L - pp : pp;
                const int pl = dir ? L - 1 - pp : pp;
                ab[k] = pp <= L ? fin[pb] : kNeg;
                al[k] = pp < L ? fin[P_pad + pl] : kNeg;
            } else {
                ab[k] = pp == 0 ? 0.f : kNeg;
                al[k] = kNeg;
            }
        }
        if (lane == 31) xchg[warp] = al[K - 1];
    }
    __syncthreads();  // mbarrier init, CSR, xchg visible

    const int C = c.chunk, NST = c.stages;
    const int64_t step_elems = (int64_t)dt * p.st;
    RingProducer prod;
    prod.src = lp_b + (int64_t)t_first * p.st;
    prod.step_elems = step_elems;
    prod.stage = 0;
    prod.remaining = nsteps;
    if (producer && lane == 0)
        for (int n = 0; n < NST; ++n) ring_issue_next(ring, prod);
    const int lagfree = GRAD ? 2 : 1;  // iterations after which a frame's slot has no reader left
    int free_at = C - 1 + lagfree;     // iteration at which the oldest in-flight stage is free

    RingPos pos;  // recursion warps: frame of step i; gradient warps: frame of step i-1
    pos.init(lp_b + (int64_t)t_first * p.st, step_elems);

    // other direction's stored rows (backward): register double buffer + L2 prefetch ahead
    float ob[K], ol[K], nb[K], nl_[K];
    double ooff = 0.0, noff = 0.0;
    const int64_t row_elems = 2 * (int64_t)P_pad;
    const float *orow = GRAD ? p.rows + ((int64_t)b * p.T + t_first) * row_elems : nullptr;
    const double *ooffp = GRAD ? p.off_rows + (int64_t)b * p.T + t_first : nullptr;
    const int64_t orow_step = (int64_t)dt * row_elems;
    auto load_other = [&](const float *row, float *vb, float *vl) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int pp = pbase + k * 32;
            const int pb = dir ? L - pp : pp;
            const int pl = dir ? L - 1 - pp : pp;
            vb[k] = pp <= L ? __ldg(row + pb) : kNeg;
            vl[k] = pp < L ? __ldg(row + P_pad + pl) : kNeg;
        }
    };
    if (GRAD && compute) {
        if (nsteps > 0) { load_other(orow, ob, ol); ooff = __ldg(ooffp); }
        if (nsteps > 1) { load_other(orow + orow_step, nb, nl_); noff = __ldg(ooffp + dt); }
    }
    float *row_out = (!GRAD && p.rows) ? p.rows + ((int64_t)b * p.T + t_first) * row_elems : nullptr;
    double *off_out = (!GRAD && p.rows) ? p.off_rows + (int64_t)b * p.T + t_first : nullptr;
    float *grow = GRAD ? p.grad + (int64_t)t_first * p.gst + (int64_t)b * p.gsb : nullptr;
    const int64_t grow_step = (int64_t)dt * p.gst;

    const int iters = GRAD ? nsteps + 1 : nsteps;
    const int gtid = tid - (W + 1) * 32, gthreads = c.G * 32;

    for (int i = 0; i < iters; ++i) {
        const int par = i & 1;
        if (producer) {
            if (i == free_at) {
                if (lane == 0) ring_issue_next(ring, prod);
                free_at += C;
            }
        } else if (compute) {
            if (i < nsteps) {
                const float *row = pos.row(ring);
                const float eb2 = fmaxf(row[p.blank] * kLog2e, kNeg);
                float el2[K];
#pragma unroll
                for (int k = 0; k < K; ++k) el2[k] = fmaxf(row[lab[k]] * kLog2e, kNeg);
                pos.advance(ring);

                float xin = warp > 0 ? xchg[par * 16 + warp - 1] : kNeg;
                if (i > 0 && (i & (kRecenter - 1)) == 0) {
                    // re-centre on the row maximum published in the previous iteration
                    float mx = wmax[0];
                    for (int w = 1; w < W; ++w) mx = fmaxf(mx, wmax[w]);
                    if (mx > kNegTest) {
#pragma unroll
                        for (int k = 0; k < K; ++k) { ab[k] -= mx; al[k] -= mx; }
                        xin -= mx;
                        off_mine += (double)mx;
                    }
                }
                // label state of the previous pair (old values): lane rotation, warp seam via smem
                float r[K];
#pragma unroll
                for (int k = 0; k < K; ++k) r[k] = __shfl_sync(FULL, al[k], (lane + 31) & 31);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float carry = lane == 0 ? (k == 0 ? xin : r[k > 0 ? k - 1 : 0]) : r[k];
                    const float A = lse2(ab[k], carry);
                    const float oth = (skipmask >> k) & 1u ? A : ab[k];
                    const float nlab = lse2(al[k], oth) + el2[k];
                    ab[k] = A + eb2;
                    al[k] = nlab;
                }
                if (lane == 31) xchg[(par ^ 1) * 16 + warp] = al[K - 1];
                if ((i & (kRecenter - 1)) == kRecenter - 1) {
                    float mx = kNeg;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int pp = pbase + k * 32;
                        if (pp <= L) mx = fmaxf(mx, ab[k]);
                        if (pp < L) mx = fmaxf(mx, al[k]);
                    }
                    mx = warp_max(mx);
                    if (lane == 0) wmax[warp] = mx;
                }

                if (!GRAD) {
                    if (row_out) {
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const int pp = pbase + k * 32;
                            if (pp <= L) row_out[dir ? L - pp : pp] = ab[k];
                            if (pp < L) row_out[P_pad + (dir ? L - 1 - pp : pp)] = al[k];
                        }
                        if (tid == 0) *off_out = off_mine;
                        row_out += orow_step;
                        off_out += dt;
                    }
                } else {
                    // posteriors of my states at this frame: 2^(alpha + beta - lp - log2 P)
                    const float bracket = (float)(off_mine + ooff + nll2);
                    float sbl = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int pp = pbase + k * 32;
                        sbl += ex2_approx(ab[k] + ob[k] - eb2 + bracket);
                        const float wl = ex2_approx(al[k] + ol[k] - el2[k] + bracket);
                        if (pp < L) wlab[par * P_pad + (dir ? L - 1 - pp : pp)] = wl;
                    }
                    const unsigned fx = __float2uint_rn(fminf(sbl, 3.5f) * 1073741824.0f);
                    const unsigned tot = __reduce_add_sync(FULL, fx);
                    if (lane == 0) atomicAdd(&blank_acc[par], tot);
                    // rotate the register prefetch and fetch two steps ahead
#pragma unroll
                    for (int k = 0; k < K; ++k) { ob[k] = nb[k]; ol[k] = nl_[k]; }
                    ooff = noff;
                    orow += orow_step;
                    ooffp += dt;
                    if (i + 2 < nsteps) {
                        load_other(orow + orow_step, nb, nl_);
                        noff = __ldg(ooffp + dt);
                    }
                    if (i + 12 < nsteps && lane == 0) {
                        const float *rowp = orow + 11 * orow_step;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const int pp = pbase + k * 32;
                            if (pp <= L) {
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + (dir ? L - pp : pp)));
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(
                                    rowp + P_pad + (dir ? max(L - 1 - pp, 0) : pp)));
                            }
                        }
                    }
                }
            }
        } else if (GRAD && i >= 1) {
            // gradient row of the frame the recursion warps finished in the previous iteration
            const int pj = (i - 1) & 1;
            const float *row = pos.row(ring);
            pos.advance(ring);
            const float *w = wlab + pj * P_pad;
            for (int cc = gtid; cc < V; cc += gthreads) {
                float rsum = 0.f;
                const int q1 = occ_start[cc + 1];
                for (int q = occ_start[cc]; q < q1; ++q) rsum += w[occ_pos[q]];
                if (cc == p.blank) {
                    rsum += (float)blank_acc[pj] * (1.0f / 1073741824.0f);
                    blank_acc[pj] = 0u;
                }
                grow[cc] = (ex2_approx(row[cc] * kLog2e) - rsum) * gs;
            }
            grow += grow_step;
        }
        __syncthreads();
    }

    if (!GRAD) {
        // frontier row for the join kernel / the backward call (natural positions)
        if (compute) {
            float *fin = p.finals + ((int64_t)b * 2 + dir) * 2 * P_pad;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int pp = pbase + k * 32;
                if (pp <= L) fin[dir ? L - pp : pp] = ab[k];
                if (pp < L) fin[P_pad + (dir ? L - 1 - pp : pp)] = al[k];
            }
            if (tid == 0) p.off_fin[(int64_t)b * 2 + dir] = off_mine;
        }
    } else if (dir == 0) {
        for (int t = Tb; t < (int)p.T; ++t) {  // frames beyond the utterance: exact zeros
            float *g = p.grad + (int64_t)t * p.gst + (int64_t)b * p.gsb;
            for (int cc = tid; cc < V; cc += blockDim.x) g[cc] = 0.f;
        }
    }
}

// Join the alpha frontier (row m-1, or the virtual start row) with the beta frontier (row m):
//   log P = off_a + off_b + lse_s( lse(alpha[s], alpha[s-1], skip ? alpha[s-2]) + beta_m[s] )
__global__ void __launch_bounds__(256) ctc_join_kernel(const CtcParams p) {
    __shared__ float red_m[8], red_s[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int L = p.tgt_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    const int P_pad = p.cfg.P_pad;
    const float *fa = p.finals + (int64_t)b * 2 * 2 * P_pad;
    const float *fb = fa + 2 * P_pad;
    const int32_t *tg = p.targets + p.tgt_off[b];
    float mx = kNeg, sm = 0.f;  // running max and sum of 2^(x - mx)
    auto push = [&](float x) {
        const float nm = fmaxf(mx, x);
        sm = sm * ex2_approx(mx - nm) + ex2_approx(x - nm);
        mx = nm;
    };
    for (int pp = tid; pp <= L; pp += 256) {
        const float a_b = fa[pp];
        const float a_lp = pp > 0 ? fa[P_pad + pp - 1] : kNeg;
        const float A = lse2(a_b, a_lp);
        push(A + fb[pp]);
        if (pp < L) {
            bool skip = false;
            if (pp > 0) skip = tg[pp] != tg[pp - 1];
            push(lse2(fa[P_pad + pp], skip ? A : a_b) + fb[P_pad + pp]);
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
    __syncthreads();
    if (tid == 0) {
        float M = red_m[0], S = red_s[0];
        for (int w = 1; w < 8; ++w) {
            const float nm = fmaxf(M, red_m[w]);
            S = S * ex2_approx(M - nm) + red_s[w] * ex2_approx(red_m[w] - nm);
            M = nm;
        }
        const bool dead = M < kNegTest;
        const double logp2 = (double)M + (double)log2f(S) + p.off_fin[(int64_t)b * 2] + p.off_fin[(int64_t)b * 2 + 1];
        p.nll[b] = dead ? __int_as_float(0x7f800000) : (float)(-logp2 * 0.6931471805599453);
        p.nll2[b] = -logp2;
    }
}

// ---------------------------------------------------------------------------- launchers
template <bool GRAD>
static int launch_lattice(const CtcParams &p, cudaStream_t stream) {
    const CtcCfg &c = p.cfg;
    const size_t smem_bytes = smem_bytes_for(c, p.V, p.Lmax, GRAD);
    if (smem_bytes > 227 * 1024) return SSAK_ERR_UNSUPPORTED;
    dim3 grid((unsigned)p.B, 2), block((c.W + 1 + (GRAD ? c.G : 0)) * 32);
#define SSAK_LAUNCH(KK)                                                                        \
    case KK: {                                                                                 \
        auto kern = ctc_lattice_kernel<KK, GRAD>;                                              \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem_bytes);                                 \
        if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }                \
        kern<<<grid, block, smem_bytes, stream>>>(p);                                          \
        break;                                                                                 \
    }
    switch (c.K) {
        SSAK_LAUNCH(1)
        SSAK_LAUNCH(2)
        SSAK_LAUNCH(4)
        SSAK_LAUNCH(8)
        default: return SSAK_ERR_UNSUPPORTED;
    }
#undef SSAK_LAUNCH
    return check_launch();
}

struct WsLayout { size_t nll2, off_fin, finals, off_rows, rows, total; };
static WsLayout ws_layout(int64_t T, int64_t B, int P_pad, bool saved) {
    WsLayout w;
    size_t o = 0;
    w.nll2 = o;     o += align_up((size_t)B * sizeof(double), 256);
    w.off_fin = o;  o += align_up((size_t)B * 2 * sizeof(double), 256);
    w.finals = o;   o += align_up((size_t)B * 2 * 2 * P_pad * sizeof(float), 256);
    w.off_rows = o; if (saved) o += align_up((size_t)B * (size_t)T * sizeof(double), 256);
    w.rows = o;     if (saved) o += align_up((size_t)B * (size_t)T * 2 * P_pad * sizeof(float), 256);
    w.total = o + 256;
    return w;
}

static int fill_params(CtcParams *p, const float *log_probs, int64_t T, int64_t B, int64_t V,
                       int64_t st, int64_t sb, const int32_t *targets, const int64_t *tgt_off,
                       const int32_t *in_len, const int32_t *tgt_len, int64_t Lmax, int32_t blank,
                       void *workspace, size_t workspace_bytes, bool saved) {
    if (!log_probs || !targets || !tgt_off || !in_len || !tgt_len || !workspace)
        return SSAK_ERR_INVALID_ARGUMENT;
    if (T < 0 || B <= 0 || V <= 0 || Lmax < 0 || blank < 0 || blank >= V || T > 0x7ffffff0 ||
        V > (1 << 20) || B > 65535 * 32)
        return SSAK_ERR_INVALID_ARGUMENT;
    if (!choose_cfg(Lmax, B, (int)V, &p->cfg)) return SSAK_ERR_UNSUPPORTED;
    const WsLayout w = ws_layout(T, B, p->cfg.P_pad, saved);
    if (workspace_bytes < w.total) return SSAK_ERR_WORKSPACE;
    p->lp = log_probs; p->T = T; p->B = B; p->V = (int)V; p->st = st; p->sb = sb;
    p->targets = targets; p->tgt_off = tgt_off; p->in_len = in_len; p->tgt_len = tgt_len;
    p->Lmax = (int)Lmax; p->blank = blank;
    char *ws = reinterpret_cast<char *>(workspace);
    p->nll2 = reinterpret_cast<double *>(ws + w.nll2);
    p->off_fin = reinterpret_cast<double *>(ws + w.off_fin);
    p->finals = reinterpret_cast<float *>(ws + w.finals);
    p->off_rows = saved ? reinterpret_cast<double *>(ws + w.off_rows) : nullptr;
    p->rows = saved ? reinterpret_cast<float *>(ws + w.rows) : nullptr;
    p->nll = nullptr; p->grad_out = nullptr; p->grad = nullptr; p->gst = p->gsb = 0;
    p->zero_inf = 0;
    return SSAK_OK;
}

}  // namespace ssak

using namespace ssak;

extern "C" size_t ssak_ctc_loss_workspace_bytes(int64_t T, int64_t B, int64_t max_target_len,
                                                int save_for_backward) {
    CtcCfg c;
    if (T < 0 || B <= 0 || max_target_len < 0 || !choose_cfg(max_target_len, B, 64, &c)) return 0;
    return ws_layout(T, B, c.P_pad, save_for_backward != 0).total;
}

extern "C" int ssak_ctc_loss_forward(const float *log_probs, int64_t T, int64_t B, int64_t V,
                                     int64_t lp_stride_t, int64_t lp_stride_b,
                                     const int32_t *targets, const int64_t *target_offsets,
                                     const int32_t *input_lengths, const int32_t *target_lengths,
                                     int64_t max_target_len, int32_t blank,
                                     int32_t save_for_backward, float *neg_log_likelihood,
                                     void *workspace, size_t workspace_bytes,
                                     ssak_stream_t stream) {
    CtcParams p;
    if (!neg_log_likelihood) return SSAK_ERR_INVALID_ARGUMENT;
    int rc = fill_params(&p, log_probs, T, B, V, lp_stride_t, lp_stride_b, targets, target_offsets,
                         input_lengths, target_lengths, max_target_len, blank, workspace,
                         workspace_bytes, save_for_backward != 0);
    if (rc != SSAK_OK) return rc;
    p.nll = neg_log_likelihood;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    rc = launch_lattice<false>(p, s);
    if (rc != SSAK_OK) return rc;
    ctc_join_kernel<<<(unsigned)B, 256, 0, s>>>(p);
    return check_launch();
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
    CtcParams p;
    if (!grad_out || !neg_log_likelihood || !grad) return SSAK_ERR_INVALID_ARGUMENT;
    int rc = fill_params(&p, log_probs, T, B, V, lp_stride_t, lp_stride_b, targets, target_offsets,
                         input_lengths, target_lengths, max_target_len, blank, workspace,
                         workspace_bytes, true);
    if (rc != SSAK_OK) return rc;
    p.nll = const_cast<float *>(neg_log_likelihood);
    p.grad_out = grad_out; p.grad = grad; p.gst = g_stride_t; p.gsb = g_stride_b;
    p.zero_inf = zero_infinity;
    return launch_lattice<true>(p, reinterpret_cast<cudaStream_t>(stream));
}
