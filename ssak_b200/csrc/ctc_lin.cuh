// ctc_lin.cuh -- CTC loss forward / backward WITHOUT a stored lattice (included by ctc_loss.cu).
//
// Same problem and same meet-in-the-middle split as the log-domain kernels of ctc_loss.cu (CTA (b,0) runs alpha over
// frames [0,m), CTA (b,1) runs beta over [m,T_b), a join kernel forms log P; the backward call continues both
// recursions over the other half, fused with the gradient), but
//   * the recursion runs in the LINEAR domain in fp64: per (blank,label) pair and frame 3 DADD/DFMA + 2 DMUL on the
//     B200's FP64 pipe (64 lanes per SM and clock) and NO transcendental -- the emissions exp(lp) are formed once per
//     (frame, used vocabulary column) by a helper warp (one MUFU.EX2 each) instead of two log-sum-exp per state;
//   * ONE warp owns the whole lattice row of its (utterance, direction): K <= 16 (blank,label) positions per lane, so
//     a frame costs one shuffle and no barrier, no cross-warp exchange, no polling (the lock-step version with 64
//     positions per warp and a named barrier per frame measured ~270 cycles per frame, all of it latency);
//   * every C = 8 frames the row is re-scaled to a maximum of ~2^400 by an exact power of two whose exponent is
//     tracked as an integer, so a state more than ~2^1470 below the row maximum is flushed to zero (fp64: 11-bit
//     exponent); see "range and the fallback" below;
//   * NOTHING of the lattice is written per frame: the forward call stores one CHECKPOINT row every C frames
//     (16 B per pair and 8 frames = 1 B per lattice cell, 8x less than the stored half lattices), and the backward
//     call recomputes the C rows of a tile from its checkpoint into a shared-memory tile private to the warp (the same
//     lane owns the same states in both directions), then runs the live direction over the tile and multiplies: posterior = live x recomputed,
//     one DMUL, because the live direction is kept in the scaling 2^(E_R) / P of the tile it is passing through.
//
// Range and the fallback.  With both directions re-scaled to their own row maximum, a state that carries a share
// gamma of the posterior mass at frame t is flushed only if  log2(alpha_max(t) beta_max(t) / P) > 1470 + log2(gamma):
// the best forward and the best backward partial path would have to disagree by a factor 2^1400 (e^970) -- garbage
// transcripts.  In the backward call the live direction overflows to inf (-> NaN gradient, loud) before that
// happens; the forward call checks the same quantity at the join frame, P == 0 for a feasible utterance, non-finite
// values, and finite log-probabilities below -700 (exp underflows fp64; SpeechBrain's -700 padding inside the
// lengths), and flags the utterance: flagged utterances are recomputed by the log-domain kernels of ctc_loss.cu
// (masked launches), so the result never depends on the range of fp64.
#pragma once

namespace ssak {
namespace lin {

constexpr int C = 8;          // frames per chunk / rows per tile / checkpoint spacing
constexpr int TARGET = 400;   // re-scaled row maximum ~ 2^TARGET
constexpr int MAXK = 16;      // positions per lane: 32 * 16 = 512 positions -> max target length 511
constexpr float kLog2eLo = 1.925963e-08f;   // log2(e) - (float)log2(e)

struct Params {
    const float *lp;
    int64_t T, B;
    int V;
    int64_t st, sb;
    const int32_t *targets;
    const int64_t *tgt_off;
    const int32_t *in_len;
    const int32_t *tgt_len;
    int Lmax, blank;
    const float *zl;      // [T][B] -log2 sum exp (logits entry points) or nullptr
    double *ck;           // [B][2][NCK][ck_row] checkpoint j = state after j*C steps (j >= 1)
    double *fr;           // [B][2][ck_row] frontier rows (state after all forward steps of the direction)
    int NCK, ck_row;      // ck_row = 2*P_pad + 2: blank states | label states | exponent | spare
    double *nll2;         // [B] -log2 P
    float *nll;           // [B]
    int *flags;           // [B] bit 0: recompute this utterance with the log-domain kernels
    int *slot;            // [B] row block of a handed-back utterance in the log-domain kernels' `rows`
    int *slot_counter;    // next free row block
    int n_slots;
    const float *grad_out;
    float *grad;
    int64_t gst, gsb;
    int zero_inf;
    int save;
    int K, G, P_pad, NST, slot_bytes, ncol_max, erow_bytes;
};

struct Smem {
    int bars, raw, ering, cid, occ_start, cursor, occ_pos, tile, wlab, blank_acc, total;
};
__host__ __device__ inline int al16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline Smem smem_map(int NST, int slot_bytes, int erow_bytes, int V, int Lmax, int K, bool grad) {
    Smem m;
    int o = 0;
    m.bars = o;      o += 8 * (4 + 4 + 2 + 2 + 2 * C + 2);    // raw_full[4] raw_empty[4] e_full[2] e_empty[2] post_full[2C] post_empty[2]
    o = (o + 127) & ~127;
    m.raw = o;       o += NST * C * slot_bytes;
    o = al16(o);
    m.ering = o;     o += 2 * C * erow_bytes;                 // two chunks: the one in use, the one being converted
    m.cid = o;       o += al16(4 * V);
    m.occ_start = o; o += al16(4 * (V + 2));
    m.cursor = o;    o += al16(4 * V);
    m.occ_pos = o;   o += al16(4 * (Lmax > 0 ? Lmax : 1));
    m.tile = o;
    if (grad) o += C * 2 * K * 32 * 8;                        // recomputed rows of the current tile (private to the recursion warp)
    m.wlab = o;
    if (grad) o += 4 * 2 * C * (32 * K + 8);
    m.blank_acc = o;
    if (grad) o += 4 * 2 * C;
    m.total = al16(o);
    return m;
}

__device__ __forceinline__ double pow2i(int e) {
    e = e < -1022 ? -1022 : (e > 1023 ? 1023 : e);
    return __hiloint2double((e + 1023) << 20, 0);
}
// x * 2^d for any |d| <= 2000 through two exact power-of-two factors (beyond: saturates -> 0 / inf, both detected)
struct Scale2 {
    double f1, f2;
};
__device__ __forceinline__ Scale2 make_scale(int d) {
    const int d1 = d < -1000 ? -1000 : (d > 1000 ? 1000 : d);
    Scale2 s;
    s.f1 = pow2i(d1);
    s.f2 = pow2i(d - d1);
    return s;
}

// exp(x) (x a natural-log probability, or a raw logit with its row normaliser zl2 in log2 units) as a double:
// 2^(x log2e + zl2) with the product carried in two floats, the fraction through MUFU.EX2 (relative error 2^-22,
// below the 1-ulp uncertainty of the fp32 input itself) and the integer part put into the fp64 exponent.
// Finite x below ~ -700 underflows fp64: `low` is set (the caller flags the utterance for the log-domain kernels).
__device__ __forceinline__ double exp_to_double(float x, float zl2, bool &low) {
    const float hi = x * kLog2e;
    const float lo = fmaf(x, kLog2e, -hi) + x * kLog2eLo;
    const float s = hi + zl2;
    const float bb = s - hi;
    const float err = (hi - (s - bb)) + (zl2 - bb);          // TwoSum: s + err == hi + zl2 (zl2 == 0: err == 0)
    const float sc = fminf(fmaxf(s, -1100.f), 1100.f);
    const float tt = sc + 12582912.f;                        // round to nearest integer without the conversion pipe
    const float nf = tt - 12582912.f;
    const int n = __float_as_int(tt) - 0x4b400000;
    const float f = (sc - nf) + (err + lo);                  // |f| <= 0.5 + tiny
    const unsigned mb = __float_as_uint(ex2_approx(f));
    double r = __hiloint2double((int)((mb >> 3) + 0x38000000u) + n * (1 << 20), (int)(mb << 29));
    if (!(s > -1000.f)) {                                    // underflow, -inf or NaN
        low = low || (x > -3.0e38f && x == x);
        r = (x != x) ? __longlong_as_double(0x7ff8000000000000ll) : 0.0;
    } else if (s > 1000.f) {
        r = __longlong_as_double(0x7ff0000000000000ll);      // (un-normalised inputs only)
    }
    return r;
}

// One frame of the recursion for the K positions of a lane.  D = 0 (alpha): position q = (blank q, label q), the
// carry is label q-1; D = 1 (beta): position q = (label q-1 [stored in l], blank q), the carry is label q [position
// q+1].  Same arithmetic both ways, only the order of the lane's positions differs:
//   A = b + carry;  t = l + b + skip * carry;  b' = A * e_blank;  l' = t * e_label.
// `visit(k, b_old, carry, A, t)` sees the position before it is overwritten (A, t: the states before their emission,
// which is what a posterior needs; b_old, carry: the aligned copy of the row the backward recomputation keeps).
template <int K, int D, typename Visit>
__device__ __forceinline__ void step(double (&b)[K], double (&l)[K], const double eb, const unsigned char *erow,
                                     const int (&eoff)[K], const double (&sk)[K], const double cin, Visit visit) {
    double carry = cin;
#pragma unroll
    for (int kk = 0; kk < K; ++kk) {
        const int k = D ? K - 1 - kk : kk;
        const double el = *reinterpret_cast<const double *>(erow + eoff[k]);
        const double A = b[k] + carry;
        const double t = fma(sk[k], carry, l[k] + b[k]);   // (a 0/1 factor: a predicated add costs ptxas a P2R/ISETP/FSEL mess)
        visit(k, b[k], carry, A, t);
        carry = l[k];
        b[k] = A * eb;
        l[k] = t * el;
    }
}

template <int K>
__device__ __forceinline__ int hi_max(const double (&b)[K], const double (&l)[K]) {
    int h = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) h = max(h, max(__double2hiint(b[k]), __double2hiint(l[k])));   // states are >= 0
    return h;
}

// One CTA per (utterance, direction).  Warp roles: 0 recursion -- the ONLY warp that touches the lattice, K positions
// per lane, so a frame needs no barrier, no cross-warp exchange: one shuffle for the lane-to-lane carry and a chain of
// three fp64 operations; 1 helper -- issues the bulk copies of the raw fp32 rows and turns the used vocabulary columns
// into the fp64 emission ring, one chunk ahead (mbarriers both ways); backward only: G gradient warps behind the
// posterior ring.
template <int K, bool GRAD, bool LOGITS>
__global__ void __launch_bounds__(GRAD ? 128 : 64) ctc_lin_kernel(const Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int b = blockIdx.x, dir = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const int G = GRAD ? p.G : 0;
    const int V = p.V, P_pad = 32 * K, NST = p.NST, slot_bytes = p.slot_bytes, erow_bytes = p.erow_bytes;

    int Tb = p.in_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.T ? (int)p.T : Tb);
    int L = p.tgt_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    const int m = Tb >> 1;
    const float *lp_b = p.lp + (int64_t)b * p.sb;
    const int32_t *tg = p.targets + p.tgt_off[b];
    const int rdir = GRAD ? 1 - dir : dir;                     // direction whose rows this launch steps through
    const int nrows = rdir ? Tb - m : m;                       // frames this CTA handles
    auto frame_of = [&](int rho) { return rdir ? Tb - 1 - rho : rho; };   // frame of local row rho of direction rdir

    const Smem sm = smem_map(NST, slot_bytes, erow_bytes, V, p.Lmax, K, GRAD);
    uint64_t *raw_full = reinterpret_cast<uint64_t *>(smem + sm.bars);
    uint64_t *raw_empty = raw_full + 4, *e_full = raw_full + 8, *e_empty = raw_full + 10;
    uint64_t *post_full = raw_full + 12, *post_empty = post_full + 2 * C;
    unsigned char *raw = smem + sm.raw, *ering = smem + sm.ering;
    int *cid = reinterpret_cast<int *>(smem + sm.cid);
    int *occ_start = reinterpret_cast<int *>(smem + sm.occ_start);
    int *cursor = reinterpret_cast<int *>(smem + sm.cursor);      // counters, then the ascending list of used columns
    int *occ_pos = reinterpret_cast<int *>(smem + sm.occ_pos);
    double *tile = reinterpret_cast<double *>(smem + sm.tile);    // [C][2K][32]
    float *wlab = reinterpret_cast<float *>(smem + sm.wlab);
    unsigned *blank_acc = reinterpret_cast<unsigned *>(smem + sm.blank_acc);
    const int WL = P_pad + 8;

    // ---- trivial outcomes of the backward call (as in ctc_lattice_kernel) ----
    float gs = 0.f;
    if (GRAD) {
        if (p.flags[b] & 1) return;                            // the log-domain kernels own this utterance
        const float nll = p.nll[b];
        gs = p.grad_out[b];
        const bool infeasible = nll == __int_as_float(0x7f800000);
        const bool isnan_ = nll != nll;
        if (infeasible || isnan_ || Tb == 0) {
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
    }

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], G > 0 ? G : 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&e_full[s], 1);
            mbar_init(&e_empty[s], 1);
        }
        if (GRAD) {
            for (int s = 0; s < 2 * C; ++s) {
                mbar_init(&post_full[s], 1);
                blank_acc[s] = 0u;
            }
            mbar_init(&post_empty[0], G);
            mbar_init(&post_empty[1], G);
        }
        mbar_fence_init();
    }
    // sentinel column of every emission-ring row: emission 0 -> the states beyond position L stay 0
    for (int i = tid; i < 2 * C; i += blockDim.x)
        *reinterpret_cast<double *>(ering + (size_t)i * erow_bytes + erow_bytes - 8) = 0.0;

    // ---- used vocabulary columns (labels of this utterance + blank), ascending, with compact ids; backward: the
    //      label-sorted order of the label states (deterministic; occ_start[c] .. occ_start[c+1] are label c's slots)
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
        if (GRAD) {
            for (int i0 = 0; i0 < L; i0 += 32) {
                const int i = i0 + lane;
                const unsigned act = __ballot_sync(FULL, i < L);
                if (i < L) {
                    int l = tg[i];
                    l = l < 0 ? 0 : (l >= V ? V - 1 : l);
                    const unsigned mm = __match_any_sync(act, l);
                    const int rank = __popc(mm & ((1u << lane) - 1u));
                    const int base = cursor[l];
                    occ_pos[i] = base + rank;
                    __syncwarp(act);
                    if (rank == 0) cursor[l] = base + __popc(mm);
                }
                __syncwarp();
            }
        }
        int np = 0;
        for (int c0 = 0; c0 < V; c0 += 32) {
            const int cc = c0 + lane;
            const bool has = cc < V && (occ_start[cc + 1] > occ_start[cc] || cc == p.blank);
            const unsigned bal = __ballot_sync(FULL, has);
            __syncwarp();
            if (has) {
                const int id = np + __popc(bal & ((1u << lane) - 1u));
                cursor[id] = cc;
                cid[cc] = id;
            }
            np += __popc(bal);
        }
        if (lane == 0) occ_start[V + 1] = np;
    }
    __syncthreads();   // last CTA-wide barrier
    const int np = occ_start[V + 1];
    const int sentinel_off = erow_bytes - 8;

    // chunk / tile geometry.  Forward: chunk n = local rows [nC, nC + C), slot f <-> row nC + f.
    // Backward: tile j = rows [jC - 1, jC + C - 1) of direction rdir (row -1 = the virtual start row), processed
    // j = J .. 0; slot i <-> row jC - 1 + i.
    const int n_seq = GRAD ? (nrows > 0 ? nrows / C + 1 : 0) : (nrows + C - 1) / C;
    auto seq_rows = [&](int n, int &row0, int &i0, int &nr) {
        if (GRAD) {
            const int j = nrows / C - n;
            row0 = j * C - 1;
            i0 = j == 0 ? 1 : 0;
            nr = nrows - row0 < C ? nrows - row0 : C;
        } else {
            row0 = n * C;
            i0 = 0;
            nr = nrows - row0 < C ? nrows - row0 : C;
        }
    };

    if (warp == 1) {
        // ================= helper: bulk copies of the raw rows + conversion into the fp64 emission ring =================
        auto issue = [&](int n) {
            int row0, i0, nr;
            seq_rows(n, row0, i0, nr);
            const int stage = n % NST;
            if (lane == 0) {
                uint32_t total = 0;
                for (int i = i0; i < nr; ++i) {
                    const uintptr_t a = reinterpret_cast<uintptr_t>(lp_b + (int64_t)frame_of(row0 + i) * p.st);
                    total += (uint32_t)(((a + 4 * V + 15) & ~(uintptr_t)15) - (a & ~(uintptr_t)15));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the stage
                mbar_arrive_expect_tx(&raw_full[stage], total);
                for (int i = i0; i < nr; ++i) {
                    const uintptr_t a = reinterpret_cast<uintptr_t>(lp_b + (int64_t)frame_of(row0 + i) * p.st);
                    const uintptr_t a0 = a & ~(uintptr_t)15;
                    bulk_g2s(raw + ((size_t)stage * C + i) * slot_bytes, reinterpret_cast<const void *>(a0),
                             (uint32_t)(((a + 4 * V + 15) & ~(uintptr_t)15) - a0), &raw_full[stage]);
                }
            }
            __syncwarp();
        };
        for (int n = 0; n < NST && n < n_seq; ++n) issue(n);
        bool low = false;
        for (int n = 0; n < n_seq; ++n) {
            const int stage = n % NST, round = n / NST;
            int row0, i0, nr;
            seq_rows(n, row0, i0, nr);
            mbar_wait(&raw_full[stage], (uint32_t)(round & 1));
            if (n >= 2) mbar_wait(&e_empty[n & 1], (uint32_t)(((n >> 1) - 1) & 1));
            // per-slot source offset and row normaliser: slot i in lane i
            int my_off = 0;
            float my_zl = 0.f;
            if (lane >= i0 && lane < nr) {
                const int t = frame_of(row0 + lane);
                const float *src = lp_b + (int64_t)t * p.st;
                my_off = (stage * C + lane) * slot_bytes + (int)(reinterpret_cast<uintptr_t>(src) & 15);
                if (LOGITS) my_zl = __ldg(p.zl + (int64_t)t * p.B + b);
            }
            unsigned char *edst = ering + (size_t)(n & 1) * C * erow_bytes;
            // row by row, two rows at a time; a lane owns the columns c = lane + 32 j of the used-column list (their
            // vocabulary indices are loop-invariant), two of them per pass: four independent conversion chains in flight
            for (int c0 = 0; c0 < np; c0 += 64) {
                const int ca = c0 + lane, cb = c0 + 32 + lane;
                const bool oka = ca < np, okb = cb < np;
                const int va = 4 * cursor[oka ? ca : 0], vb = 4 * cursor[okb ? cb : 0];
                for (int i = i0; i < nr; i += 2) {
                    const bool two = i + 1 < nr;
                    const int off0 = __shfl_sync(FULL, my_off, i), off1 = __shfl_sync(FULL, my_off, two ? i + 1 : i);
                    const float z0 = __shfl_sync(FULL, my_zl, i), z1 = __shfl_sync(FULL, my_zl, two ? i + 1 : i);
                    const float x00 = *reinterpret_cast<const float *>(raw + off0 + va);
                    const float x01 = *reinterpret_cast<const float *>(raw + off0 + vb);
                    const float x10 = *reinterpret_cast<const float *>(raw + off1 + va);
                    const float x11 = *reinterpret_cast<const float *>(raw + off1 + vb);
                    const double r00 = exp_to_double(x00, z0, low), r01 = exp_to_double(x01, z0, low);
                    const double r10 = exp_to_double(x10, z1, low), r11 = exp_to_double(x11, z1, low);
                    double *d0 = reinterpret_cast<double *>(edst + (size_t)i * erow_bytes);
                    double *d1 = reinterpret_cast<double *>(edst + (size_t)(i + 1) * erow_bytes);
                    if (oka) d0[ca] = r00;
                    if (okb) d0[cb] = r01;
                    if (two && oka) d1[ca] = r10;
                    if (two && okb) d1[cb] = r11;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&e_full[n & 1]);
            if (n + NST < n_seq) {
                if (GRAD) mbar_wait(&raw_empty[stage], (uint32_t)(round & 1));   // the gradient warps read the raw rows too
                issue(n + NST);
            }
        }
        if (__any_sync(FULL, low) && lane == 0) atomicOr(&p.flags[b], 1);
        return;
    }

    if (GRAD && warp >= 2) {
        // ================= gradient warps: consume the posterior ring (as in ctc_lattice_kernel) =================
        const int gwarp = warp - 2;
        const bool vec_ok = (V & 3) == 0 && ((p.gst | p.gsb) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.grad) & 15) == 0;
        unsigned present = 0;
        {
            int j = 0;
            for (int cc = lane; cc < V; cc += 32, ++j)
                if (j < 32 && (occ_start[cc + 1] > occ_start[cc] || cc == p.blank)) present |= 1u << j;
        }
        auto label_mass = [&](int cc, const float *w, int slot) {
            float rsum = 0.f;
            const int q1 = occ_start[cc + 1];
            for (int q = occ_start[cc]; q < q1; ++q) rsum += w[q];
            if (cc == p.blank) rsum += (float)blank_acc[slot] * (1.0f / 1073741824.0f);
            return rsum;
        };
        for (int n = 0; n < n_seq; ++n) {
            const int stage = n % NST, round = n / NST;
            int row0, i0, nr;
            seq_rows(n, row0, i0, nr);
            mbar_wait(&raw_full[stage], (uint32_t)(round & 1));
            const int pbuf = (n & 1) * C;
            for (int i = C - 1 - gwarp; i >= 0; i -= G) {     // slots in the order the recursion produces them
                const int slot = pbuf + i;
                mbar_wait(&post_full[slot], (uint32_t)((n >> 1) & 1));
                if (i < i0 || i >= nr) continue;
                const int t = frame_of(row0 + i);
                const float *src = lp_b + (int64_t)t * p.st;
                const unsigned char *rowb = raw + ((size_t)stage * C + i) * slot_bytes + (reinterpret_cast<uintptr_t>(src) & 15);
                const float zl = LOGITS ? __ldg(p.zl + (int64_t)t * p.B + b) : 0.f;
                const float *w = wlab + slot * WL;
                float *grow = p.grad + (int64_t)t * p.gst + (int64_t)b * p.gsb;
                if (vec_ok && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
                    const float4 *row4 = reinterpret_cast<const float4 *>(rowb);
                    float4 *g4 = reinterpret_cast<float4 *>(grow);
                    for (int it = lane; it < (V >> 2); it += 32) {
                        const float4 x = row4[it];
                        g4[it] = make_float4(ex2_approx(fmaf(x.x, kLog2e, zl)) * gs, ex2_approx(fmaf(x.y, kLog2e, zl)) * gs,
                                             ex2_approx(fmaf(x.z, kLog2e, zl)) * gs, ex2_approx(fmaf(x.w, kLog2e, zl)) * gs);
                    }
                    __syncwarp();
                    const float *row = reinterpret_cast<const float *>(rowb);
                    for (int ii = lane; ii < np; ii += 32) {
                        const int cc = cursor[ii];
                        grow[cc] = (ex2_approx(fmaf(row[cc], kLog2e, zl)) - label_mass(cc, w, slot)) * gs;
                    }
                } else {
                    const float *row = reinterpret_cast<const float *>(rowb);
                    int j = 0;
                    for (int cc = lane; cc < V; cc += 32, ++j) {
                        float val = ex2_approx(fmaf(row[cc], kLog2e, zl));
                        if (j >= 32 || ((present >> j) & 1u)) val -= label_mass(cc, w, slot);
                        grow[cc] = val * gs;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&raw_empty[stage]);
                mbar_arrive(&post_empty[n & 1]);
            }
        }
        if (dir == 0) {   // frames beyond the utterance: exact zeros
            for (int t = Tb + gwarp; t < (int)p.T; t += G) {
                float *g = p.grad + (int64_t)t * p.gst + (int64_t)b * p.gsb;
                for (int cc = lane; cc < V; cc += 32) g[cc] = 0.f;
            }
        }
        return;
    }

    // ================= recursion warp =================
    const int q0 = lane * K;                                   // my first position
    const int eoff_blank = 8 * cid[p.blank];
    // static tables of a direction D at my positions: emission offset of the label state, skip bit
    auto tables = [&](int D, int (&eoff)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int q = q0 + k, li = D ? q - 1 : q;
            eoff[k] = sentinel_off;
            if (q <= L && li >= 0 && li < L) {
                int l = tg[li];
                l = l < 0 ? 0 : (l >= V ? V - 1 : l);
                eoff[k] = 8 * cid[l];
            }
        }
    };
    // skip factor of position q: labels q-1 and q both exist and differ -- the SAME condition for alpha (label q may
    // be entered from label q-1) and for beta (label q-1 may continue into label q): one table for both directions
    double sk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int q = q0 + k;
        sk[k] = 0.0;
        if (q >= 1 && q < L) {
            int l1 = tg[q - 1], l2 = tg[q];
            l1 = l1 < 0 ? 0 : (l1 >= V ? V - 1 : l1);
            l2 = l2 < 0 ? 0 : (l2 >= V ? V - 1 : l2);
            if (l1 != l2) sk[k] = 1.0;
        }
    }
    auto load_row = [&](const double *row, double (&bb)[K], double (&ll)[K]) -> int {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int q = q0 + k;
            bb[k] = q <= L ? row[q] : 0.0;
            ll[k] = q <= L ? row[P_pad + q] : 0.0;
        }
        return (int)row[2 * P_pad];
    };
    auto store_row = [&](double *row, const double (&bb)[K], const double (&ll)[K], int E) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int q = q0 + k;
            if (q <= L) {
                row[q] = bb[k];
                row[P_pad + q] = ll[k];
            }
        }
        if (lane == 0) row[2 * P_pad] = (double)E;
    };
    auto start_row = [&](int D, double (&bb)[K], double (&ll)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            bb[k] = (q0 + k == (D ? L : 0)) ? 1.0 : 0.0;
            ll[k] = 0.0;
        }
    };
    // lane-to-lane carry of direction D: the boundary label state of the neighbour lane, 0 at the end of the chain
    auto carry_in = [&](int D, const double (&ll)[K]) -> double {
        const double nb = D ? __shfl_down_sync(FULL, ll[0], 1) : __shfl_up_sync(FULL, ll[K - 1], 1);
        return lane == (D ? 31 : 0) ? 0.0 : nb;
    };
    double *ck_base = p.ck + ((int64_t)b * 2) * p.NCK * p.ck_row;
    double *fr_base = p.fr + ((int64_t)b * 2) * p.ck_row;

    if (!GRAD) {
        // ---------------- forward: one direction, checkpoint every C rows, frontier at the end ----------------
        int eoff[K];
        double bs[K], ls[K];
        tables(dir, eoff);
        start_row(dir, bs, ls);
        int E = 0;
        double *ck_dir = ck_base + (int64_t)dir * p.NCK * p.ck_row;
        auto nop = [](int, double, double, double, double) {};
        for (int n = 0; n < n_seq; ++n) {
            int row0, i0, nr;
            seq_rows(n, row0, i0, nr);
            mbar_wait(&e_full[n & 1], (uint32_t)((n >> 1) & 1));
            const unsigned char *er = ering + (size_t)(n & 1) * C * erow_bytes;
            auto frame = [&](const int f) {
                const unsigned char *erow = er + f * erow_bytes;
                const double eb = *reinterpret_cast<const double *>(erow + eoff_blank);
                if (dir) step<K, 1>(bs, ls, eb, erow, eoff, sk, carry_in(1, ls), nop);
                else step<K, 0>(bs, ls, eb, erow, eoff, sk, carry_in(0, ls), nop);
            };
            // (not unrolled over the frames: a frame is already 13 x ~10 instructions, and eight copies of it cost
            //  instruction-cache misses -- "no_instruction" was the second stall reason of the unrolled version)
#pragma unroll 1
            for (int f = 0; f < nr; ++f) frame(f);
            __syncwarp();
            if (lane == 0) mbar_arrive(&e_empty[n & 1]);
            // re-scale the row to a maximum of ~2^TARGET (the warp owns the whole row: one integer REDUX)
            const int M = __reduce_max_sync(FULL, hi_max<K>(bs, ls));
            int d = 0;
            if (M >= 0x7ff00000) {
                if (lane == 0) atomicOr(&p.flags[b], 1);        // inf / NaN: let the log-domain kernels decide
            } else if (M >= 0x00100000) {
                d = TARGET - ((M >> 20) - 1023);
            }
            if (d != 0) {
                const Scale2 sc = make_scale(d);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    bs[k] = bs[k] * sc.f1 * sc.f2;
                    ls[k] = ls[k] * sc.f1 * sc.f2;
                }
                E -= d;
            }
            if (p.save && nr == C) store_row(ck_dir + (int64_t)(n + 1) * p.ck_row, bs, ls, E);
        }
        store_row(fr_base + (int64_t)dir * p.ck_row, bs, ls, E);
        return;
    }

    // ---------------- backward: recompute direction DR = 1 - dir tile by tile, run direction dir live ----------------
    if (nrows == 0) return;
    const int DL = dir, DR = 1 - dir;
    int eoffL[K], eoffR[K], wl_off[K];
    double lb[K], ll[K];
    tables(DL, eoffL);
    tables(DR, eoffR);
#pragma unroll
    for (int k = 0; k < K; ++k) {   // slot of my live label state in the label-sorted posterior buffer
        const int q = q0 + k, li = DL ? q - 1 : q;
        wl_off[k] = (q <= L && li >= 0 && li < L) ? occ_pos[li] : WL - 1;
    }
    const int EL0 = load_row(fr_base + (int64_t)DL * p.ck_row, lb, ll);
    const double log2P = -p.nll2[b];
    const double Epd = floor(log2P);
    const double invPm = 1.0 / exp2(log2P - Epd);
    const int Ep = (int)Epd;
    int prevER = 0;
    const double *ck_dir = ck_base + (int64_t)DR * p.NCK * p.ck_row;
    double *tl = tile + lane;                                   // my column of the tile: entry (i, kk) at (i * 2K + kk) * 32
    bool bad = false;
    for (int n = 0; n < n_seq; ++n) {
        int row0, i0, nr;
        seq_rows(n, row0, i0, nr);
        const int j = nrows / C - n;
        mbar_wait(&e_full[n & 1], (uint32_t)((n >> 1) & 1));
        const unsigned char *er = ering + (size_t)(n & 1) * C * erow_bytes;
        // ---- R phase: rows row0 .. row0 + nr - 1 of direction DR from checkpoint j (or the virtual start row);
        //      the aligned copy of every row (blank state, and the label state the other direction pairs it with =
        //      the carry this position receives) goes to my column of the tile
        {
            double rb[K], rl[K];
            int ER = 0;
            if (j > 0) ER = load_row(ck_dir + (int64_t)j * p.ck_row, rb, rl); else start_row(DR, rb, rl);
            // the live direction moves into this tile's scaling: live_hat = live_true * 2^(ER) / P
            {
                const Scale2 sc = make_scale(n == 0 ? EL0 + ER - Ep : ER - prevER);
                const double f1 = n == 0 ? sc.f1 * invPm : sc.f1;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    lb[k] = lb[k] * f1 * sc.f2;
                    ll[k] = ll[k] * f1 * sc.f2;
                }
                prevER = ER;
            }
#pragma unroll 1
            for (int i = 0; i < C; ++i) {
                if (i < nr) {
                    const double cin = carry_in(DR, rl);
                    double *ti = tl + (size_t)i * 2 * K * 32;
                    if (i + 1 < nr) {
                        const unsigned char *erow = er + (i + 1) * erow_bytes;
                        const double eb = *reinterpret_cast<const double *>(erow + eoff_blank);
                        auto keep = [&](int k, double b_old, double carry, double, double) {
                            ti[k * 32] = b_old;
                            ti[(K + k) * 32] = carry;
                        };
                        if (DR) step<K, 1>(rb, rl, eb, erow, eoffR, sk, cin, keep);
                        else step<K, 0>(rb, rl, eb, erow, eoffR, sk, cin, keep);
                    } else {
                        // last row of the tile: only its aligned copy (the carries the next step would have fetched)
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            ti[k * 32] = rb[k];
                            const double c = DR ? (k == K - 1 ? cin : rl[k + 1 < K ? k + 1 : k]) : (k == 0 ? cin : rl[k > 0 ? k - 1 : 0]);
                            ti[(K + k) * 32] = c;
                        }
                    }
                }
            }
        }
        // ---- B phase: the live direction over the tile's rows, last row first
        if (n >= 2) mbar_wait(&post_empty[n & 1], (uint32_t)(((n >> 1) - 1) & 1));
        const int pbuf = (n & 1) * C;
#pragma unroll 1
        for (int i = C - 1; i >= 0; --i) {
            if (i >= i0 && i < nr) {
                const unsigned char *erow = er + i * erow_bytes;
                const double eb = *reinterpret_cast<const double *>(erow + eoff_blank);
                const double *ti = tl + (size_t)i * 2 * K * 32;
                float *wl = wlab + (pbuf + i) * WL;
                float sbl = 0.f;
                // posteriors of my states at this frame: (live state before its emission) x (recomputed state)
                auto post = [&](int k, double, double, double A, double t) {
                    sbl += (float)(A * ti[k * 32]);
                    wl[wl_off[k]] = (float)(t * ti[(K + k) * 32]);
                };
                const double cin = carry_in(DL, ll);
                if (DL) step<K, 1>(lb, ll, eb, erow, eoffL, sk, cin, post);
                else step<K, 0>(lb, ll, eb, erow, eoffL, sk, cin, post);
                const unsigned fx = __float2uint_rn(fminf(sbl, 3.5f) * 1073741824.0f);
                const unsigned tot = __reduce_add_sync(FULL, fx);
                bad = bad || !(sbl <= 3.0e38f);                 // inf / NaN: the live direction left the range of fp64
                __syncwarp();
                if (lane == 0) {
                    blank_acc[pbuf + i] = tot;
                    mbar_arrive(&post_full[pbuf + i]);
                }
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&post_full[pbuf + i]);   // unused slot of a partial tile: keep the phases in step
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&e_empty[n & 1]);
    }
    if (__any_sync(FULL, bad) && lane == 0) atomicOr(&p.flags[b], 2);
}

// log P = log2( sum over the transitions from the alpha frontier (row m-1) into the beta frontier (row m) ) + exponents;
// feasibility (T_b >= L_b + repeats) decides between +inf and "let the log-domain kernels look at it".
__global__ void __launch_bounds__(256) ctc_lin_join_kernel(const Params p) {
    __shared__ double red[8];
    __shared__ int redh[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int L = p.tgt_len[b];
    L = L < 0 ? 0 : (L > p.Lmax ? p.Lmax : L);
    int Tb = p.in_len[b];
    Tb = Tb < 0 ? 0 : (Tb > (int)p.T ? (int)p.T : Tb);
    const int P_pad = p.P_pad;
    const double *fa = p.fr + ((int64_t)b * 2) * p.ck_row, *fb = fa + p.ck_row;
    const int32_t *tg = p.targets + p.tgt_off[b];
    double sum = 0.0;
    int rep = 0, badt = 0, ha = 0, hb = 0;
    for (int q = tid; q <= L; q += 256) {
        const double ab = fa[q];
        const double alp = q > 0 ? fa[P_pad + q - 1] : 0.0;
        const double A = ab + alp;
        sum += A * fb[q];
        ha = max(ha, __double2hiint(ab));
        hb = max(hb, __double2hiint(fb[q]));
        if (q < L) {
            const bool skip = q > 0 && tg[q] != tg[q - 1];
            const double al = fa[P_pad + q];
            const double t = al + ab + (skip ? alp : 0.0);
            const double bl = fb[P_pad + q + 1];               // beta's label q sits at position q + 1
            sum += t * bl;
            rep += (q > 0 && tg[q] == tg[q - 1]) ? 1 : 0;
            badt |= (tg[q] < 0 || tg[q] >= p.V) ? 1 : 0;
            ha = max(ha, __double2hiint(al));
            hb = max(hb, __double2hiint(bl));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        rep += __shfl_xor_sync(0xffffffffu, rep, d);
        ha = max(ha, __shfl_xor_sync(0xffffffffu, ha, d));
        hb = max(hb, __shfl_xor_sync(0xffffffffu, hb, d));
    }
    if (lane == 0) {
        red[warp] = sum;
        redh[warp] = rep;
    }
    badt = __syncthreads_or(badt);
    __shared__ int sha[8], shb[8];
    if (lane == 0) {
        sha[warp] = ha;
        shb[warp] = hb;
    }
    __syncthreads();
    if (tid == 0) {
        double S = 0.0;
        int R = 0, HA = 0, HB = 0;
        for (int w = 0; w < 8; ++w) {
            S += red[w];
            R += redh[w];
            HA = max(HA, sha[w]);
            HB = max(HB, shb[w]);
        }
        const double Ea = fa[2 * P_pad], Eb = fb[2 * P_pad];
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
            bool hand_back = (p.flags[b] & 1) != 0;             // set by the forward kernel (emission range, inf / NaN)
            if (!(S > 0.0) || !(S < 1.0e300)) {
                hand_back = true;                               // zero, inf or NaN: the log-domain kernels decide
                p.nll[b] = QNAN;
                p.nll2[b] = 0.0;
            } else {
                const double l2 = log2(S);
                // mismatch of the two frontiers, log2(alpha_max beta_max / P) at the join frame
                const double mism = (double)(((HA >> 20) - 1023) + ((HB >> 20) - 1023)) - l2;
                if (mism > 1000.0) hand_back = true;
                const double logp2 = l2 + Ea + Eb;
                p.nll[b] = (float)(-logp2 * 0.6931471805599453);
                p.nll2[b] = -logp2;
            }
            if (hand_back) {
                // a row block for the log-domain kernels' stored half lattices; none left: NaN likelihood (loud)
                const int sl = atomicAdd(p.slot_counter, 1);
                if (sl < p.n_slots) {
                    p.slot[b] = sl;
                    p.flags[b] = 1;
                } else {
                    p.flags[b] = 4;                             // neither path owns it: the backward fills NaN
                    p.nll[b] = QNAN;
                }
            }
        }
    }
}

}  // namespace lin
}  // namespace ssak
