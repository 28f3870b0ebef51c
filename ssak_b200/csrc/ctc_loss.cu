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
//     cp.async.bulk (1-D TMA, completion on an mbarrier) and gathered at the label columns.
//   * Backward: dedicated gradient warps run one frame behind the recursion warps.  The
//     recursion warps publish per-state posteriors (shared memory), blank posteriors are
//     summed with an integer warp reduction in 2^-30 fixed point (deterministic), label
//     posteriors are summed by the gradient warps through a per-utterance CSR
//     (label -> positions) built once, and full gradient rows are written coalesced.
#include "common.cuh"

namespace ssak {

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
    float *rows;    // [B][T][2*P_pad] half lattices (nullptr: not saved)
    float *finals;  // [B][2][2*P_pad] frontier rows
    float *nll2;    // [B] -log2 P kept in the workspace (no ln2 round trip before the backward)
    float *nll;     // [B]
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
    // regime): fat lanes, few warps, so several utterances share an SM without barriers
    // between many warps.
    int wtarget = (2 * B <= 2 * 148) ? 8 : ((2 * B <= 6 * 148) ? 4 : 2);
    wtarget = env_int("SSAK_CTC_WARPS", wtarget);
    int K = env_int("SSAK_CTC_K", 0);
    if (K == 0) {
        K = 1;
        while (K < 16 && (P + 32 * K - 1) / (32 * K) > wtarget) K *= 2;
    }
    if (K != 1 && K != 2 && K != 4 && K != 8 && K != 16) return false;
    const int64_t W = (P + 32 * K - 1) / (32 * K);
    if (W > 16) return false;  // L <= 8191
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

struct SmemLayout {
    size_t full, xchg, blank_acc, ring, wlab, occ_start, cursor, occ_pos, total;
};
static SmemLayout smem_layout(const CtcCfg &c, int V, int Lmax, bool grad) {
    SmemLayout s;
    size_t o = 0;
    s.full = o;      o += 8 * 8;                       // up to 8 stages
    s.xchg = o;      o += 2 * 16 * sizeof(float);
    s.blank_acc = o; o += 16;
    o = align_up(o, 16);
    s.ring = o;      o += (size_t)c.stages * c.chunk * c.slot_bytes;
    s.wlab = o;
    if (grad) {
        o += 2 * (size_t)c.P_pad * sizeof(float);
        s.occ_start = o; o += ((size_t)V + 1) * sizeof(int);
        s.cursor = o;    o += (size_t)V * sizeof(int);
        s.occ_pos = o;   o += (size_t)(Lmax > 0 ? Lmax : 1) * sizeof(int);
    } else {
        s.occ_start = s.cursor = s.occ_pos = o;
    }
    s.total = align_up(o, 16);
    return s;
}

// ------------------------------------------------------------------------------ kernel
template <int K, bool GRAD>
__global__ void __launch_bounds__(GRAD ? 640 : 512, 1) ctc_lattice_kernel(const CtcParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const CtcCfg &c = p.cfg;
    const int b = blockIdx.x;
    const int dir = blockIdx.y;  // 0: alpha (forward in time), 1: beta (backward in time)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool compute = warp < c.W;
    const unsigned FULL = 0xffffffffu;

    int Tb = p.in_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.T ? (int)p.T : Tb);
    int L = p.tgt_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    const int m = Tb >> 1;
    const int n1 = dir ? Tb - m : m;                 // frames this direction owns in forward()
    const int tau0 = GRAD ? n1 : 0;                  // first direction-local step of this launch
    const int nsteps = GRAD ? Tb - n1 : n1;
    const int P_pad = c.P_pad;
    const int V = p.V;
    const float *lp_b = p.lp + (int64_t)b * p.sb;
    const int32_t *tg = p.targets + p.tgt_off[b];

    // ---- shared memory carve-up (same layout function as the host) ----
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    float *xchg = reinterpret_cast<float *>(smem + 64);                  // [2][16]
    unsigned *blank_acc = reinterpret_cast<unsigned *>(smem + 64 + 128); // [2]
    size_t off = (64 + 128 + 16 + 15) & ~(size_t)15;
    RowRing ring;
    ring.slots = smem + off;
    ring.full = full;
    ring.chunk = c.chunk;
    ring.stages = c.stages;
    ring.slot_bytes = c.slot_bytes;
    ring.row_bytes = 4 * V;
    off += (size_t)c.stages * c.chunk * c.slot_bytes;
    float *wlab = reinterpret_cast<float *>(smem + off);                 // [2][P_pad]  (GRAD)
    int *occ_start = reinterpret_cast<int *>(smem + off + 2 * (size_t)P_pad * sizeof(float));
    int *cursor = occ_start + (V + 1);
    int *occ_pos = cursor + V;

    // ---- gradient prologue: trivial outcomes ----
    float nll2 = 0.f, gs = 0.f;
    if (GRAD) {
        const float nll = p.nll[b];
        gs = p.grad_out[b];
        const bool infeasible = !(nll < 3.0e38f);    // +inf (or NaN)
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
            const int pp = pbase + k * 32;            // direction-local pair index
            int l = p.blank, lprev = -1;
            if (pp < L) {
                const int li = dir ? L - 1 - pp : pp; // natural label index
                l = tg[li];
                l = l < 0 ? 0 : (l >= V ? V - 1 : l);
                if (pp >= 1) {
                    lprev = tg[dir ? li + 1 : li - 1];
                    lprev = lprev < 0 ? 0 : (lprev >= V ? V - 1 : lprev);
                    if (lprev != l) skipmask |= 1u << k;
                }
            }
            lab[k] = l;
        }
    }

    // ---- CSR label -> natural positions (backward only) ----
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
    if (compute) {
        const float *fin = p.finals + ((int64_t)b * 2 + dir) * 2 * P_pad;
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
    const int nchunks = (nsteps + C - 1) / C;
    const int dt = dir ? -1 : 1;
    auto frame_of = [&](int i) { return dir ? Tb - 1 - (tau0 + i) : tau0 + i; };
    if (tid == 0) {
        for (int n = 0; n < NST && n < nchunks; ++n) {
            const int cnt = min(C, nsteps - n * C);
            ring_issue(ring, n, lp_b, p.st, frame_of(n * C), dt, cnt);
        }
    }

    // other direction's stored rows (backward): register double buffer + L2 prefetch ahead
    float ob[K], ol[K], nb[K], nl_[K];
    const float *rows_b = GRAD ? p.rows + (int64_t)b * p.T * 2 * P_pad : nullptr;
    auto load_other = [&](int i, float *vb, float *vl) {
        const float *row = rows_b + (int64_t)frame_of(i) * 2 * P_pad;
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
        if (nsteps > 0) load_other(0, ob, ol);
        if (nsteps > 1) load_other(1, nb, nl_);
    }

    const int lagfree = GRAD ? 2 : 1;
    const int iters = GRAD ? nsteps + 1 : nsteps;
    const int gtid = tid - c.W * 32, gthreads = c.G * 32;

    for (int i = 0; i < iters; ++i) {
        const int par = i & 1;
        // refill the ring stage whose last reader finished before the previous barrier
        if (tid == 0) {
            const int j = i - lagfree;
            if (j >= 0 && (j % C) == C - 1) {
                const int nxt = j / C + NST;
                if (nxt < nchunks) {
                    const int cnt = min(C, nsteps - nxt * C);
                    ring_issue(ring, nxt % NST, lp_b, p.st, frame_of(nxt * C), dt, cnt);
                }
            }
        }
        if (compute && i < nsteps) {
            const int n = i / C, f = i - n * C, stage = n % NST;
            if (f == 0) mbar_wait(&full[stage], (n / NST) & 1);
            const int t = frame_of(i);
            const float *row = ring_row(ring, stage, f, lp_b, p.st, t);
            const float eb2 = fmaxf(row[p.blank] * kLog2e, kNeg);
            float el2[K];
#pragma unroll
            for (int k = 0; k < K; ++k) el2[k] = fmaxf(row[lab[k]] * kLog2e, kNeg);

            // label state of the previous pair (old values): lane rotation, warp seam via smem
            float r[K];
#pragma unroll
            for (int k = 0; k < K; ++k) r[k] = __shfl_sync(FULL, al[k], (lane + 31) & 31);
            const float xin = warp > 0 ? xchg[par * 16 + warp - 1] : kNeg;
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

            if (!GRAD) {
                if (p.rows) {
                    float *row_o = p.rows + ((int64_t)b * p.T + t) * 2 * P_pad;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int pp = pbase + k * 32;
                        if (pp <= L) row_o[dir ? L - pp : pp] = ab[k];
                        if (pp < L) row_o[P_pad + (dir ? L - 1 - pp : pp)] = al[k];
                    }
                }
            } else {
                // posteriors of my states at frame t: 2^(alpha + beta - lp + nll)
                float sbl = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int pp = pbase + k * 32;
                    sbl += ex2_approx(ab[k] + ob[k] - eb2 + nll2);
                    const float wl = ex2_approx(al[k] + ol[k] - el2[k] + nll2);
                    if (pp < L) wlab[par * P_pad + (dir ? L - 1 - pp : pp)] = wl;
                }
                const unsigned fx = __float2uint_rn(fminf(sbl, 3.5f) * 1073741824.0f);
                const unsigned tot = __reduce_add_sync(FULL, fx);
                if (lane == 0) atomicAdd(&blank_acc[par], tot);
                // rotate the register prefetch and fetch two steps ahead
#pragma unroll
                for (int k = 0; k < K; ++k) { ob[k] = nb[k]; ol[k] = nl_[k]; }
                if (i + 2 < nsteps) load_other(i + 2, nb, nl_);
                if (i + 10 < nsteps && lane == 0) {
                    const float *rowp = rows_b + (int64_t)frame_of(i + 10) * 2 * P_pad;
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
        if (GRAD && !compute && i >= 1) {
            // gradient row of the frame the recursion warps finished in the previous iteration
            const int j = i - 1, pj = j & 1;
            const int n = j / C, f = j - n * C, stage = n % NST;
            if (f == 0) mbar_wait(&full[stage], (n / NST) & 1);
            const int t = frame_of(j);
            const float *row = ring_row(ring, stage, f, lp_b, p.st, t);
            float *g = p.grad + (int64_t)t * p.gst + (int64_t)b * p.gsb;
            const float *w = wlab + pj * P_pad;
            for (int cc = gtid; cc < V; cc += gthreads) {
                float rsum = 0.f;
                const int q1 = occ_start[cc + 1];
                for (int q = occ_start[cc]; q < q1; ++q) rsum += w[occ_pos[q]];
                if (cc == p.blank) {
                    rsum += (float)blank_acc[pj] * (1.0f / 1073741824.0f);
                    blank_acc[pj] = 0u;
                }
                g[cc] = (ex2_approx(row[cc] * kLog2e) - rsum) * gs;
            }
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
        }
    } else if (dir == 0) {
        for (int t = Tb; t < (int)p.T; ++t) {  // frames beyond the utterance: exact zeros
            float *g = p.grad + (int64_t)t * p.gst + (int64_t)b * p.gsb;
            for (int cc = tid; cc < V; cc += blockDim.x) g[cc] = 0.f;
        }
    }
}

// Join the alpha frontier (row m-1, or the virtual start row) with the beta frontier (row m):
//   log P = lse_s( lse(alpha[s], alpha[s-1], skip ? alpha[s-2]) + beta_m[s] )
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
        const float logp2 = M + lg2_approx(S);
        p.nll[b] = (logp2 < kNegTest) ? __int_as_float(0x7f800000) : -logp2 * kLn2;
        p.nll2[b] = -logp2;
    }
}

// ---------------------------------------------------------------------------- launchers
template <bool GRAD>
static int launch_lattice(const CtcParams &p, cudaStream_t stream) {
    const CtcCfg &c = p.cfg;
    const SmemLayout sl = smem_layout(c, p.V, p.Lmax, GRAD);
    if (sl.total > 227 * 1024) return SSAK_ERR_UNSUPPORTED;
    dim3 grid((unsigned)p.B, 2), block((c.W + (GRAD ? c.G : 0)) * 32);
    if ((int)block.x > (GRAD ? 640 : 512)) return SSAK_ERR_UNSUPPORTED;
#define SSAK_LAUNCH(KK)                                                                        \
    case KK: {                                                                                 \
        auto kern = ctc_lattice_kernel<KK, GRAD>;                                              \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)sl.total);                                   \
        if (e != cudaSuccess) { set_last_cuda_error(e); return SSAK_ERR_CUDA; }                \
        kern<<<grid, block, sl.total, stream>>>(p);                                            \
        break;                                                                                 \
    }
    switch (c.K) {
        SSAK_LAUNCH(1)
        SSAK_LAUNCH(2)
        SSAK_LAUNCH(4)
        SSAK_LAUNCH(8)
        SSAK_LAUNCH(16)
        default: return SSAK_ERR_UNSUPPORTED;
    }
#undef SSAK_LAUNCH
    return check_launch();
}

static int fill_params(CtcParams *p, const float *log_probs, int64_t T, int64_t B, int64_t V,
                       int64_t st, int64_t sb, const int32_t *targets, const int64_t *tgt_off,
                       const int32_t *in_len, const int32_t *tgt_len, int64_t Lmax, int32_t blank,
                       void *workspace, size_t workspace_bytes, bool saved) {
    if (!log_probs || !targets || !tgt_off || !in_len || !tgt_len || !workspace)
        return SSAK_ERR_INVALID_ARGUMENT;
    if (T < 0 || B <= 0 || V <= 0 || Lmax < 0 || blank < 0 || blank >= V || T > 0x7ffffff0 ||
        V > (1 << 20))
        return SSAK_ERR_INVALID_ARGUMENT;
    if (!choose_cfg(Lmax, B, (int)V, &p->cfg)) return SSAK_ERR_UNSUPPORTED;
    if (workspace_bytes < ssak_ctc_loss_workspace_bytes(T, B, Lmax, saved ? 1 : 0))
        return SSAK_ERR_WORKSPACE;
    p->lp = log_probs; p->T = T; p->B = B; p->V = (int)V; p->st = st; p->sb = sb;
    p->targets = targets; p->tgt_off = tgt_off; p->in_len = in_len; p->tgt_len = tgt_len;
    p->Lmax = (int)Lmax; p->blank = blank;
    char *ws = reinterpret_cast<char *>(workspace);
    const size_t hdr_bytes = align_up((size_t)B * sizeof(float), 256);
    const size_t fin_bytes = align_up((size_t)B * 2 * 2 * p->cfg.P_pad * sizeof(float), 256);
    p->nll2 = reinterpret_cast<float *>(ws);
    p->finals = reinterpret_cast<float *>(ws + hdr_bytes);
    p->rows = saved ? reinterpret_cast<float *>(ws + hdr_bytes + fin_bytes) : nullptr;
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
    size_t bytes = align_up((size_t)B * sizeof(float), 256) +
                   align_up((size_t)B * 2 * 2 * c.P_pad * sizeof(float), 256);
    if (save_for_backward) bytes += (size_t)B * (size_t)T * 2 * c.P_pad * sizeof(float);
    return bytes + 256;
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
