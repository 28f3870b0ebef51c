// ctc_lin32_lv.cuh -- the throughput loss kernels for LARGE vocabularies (128 < V <= 2048, V % 4 == 0, 16-byte aligned
// rows, targets up to 223 labels: BPE / word-piece CTC heads, BASELINE config C5).  Included by ctc_lin32.cu; same
// chains, same block floating point, same checkpoints / tile / self-check as the kernels there.  What differs is how
// emissions reach the recursion:
//   * whole rows travel global -> shared with cp.async.bulk (1-D TMA: one instruction per 4 KB row, completion on a
//     per-warp mbarrier) into a per-warp ring of rows; every lane then picks the <= K + 2 columns its own positions
//     use (K + 1 labels -- the one table that serves both directions -- and blank) out of shared memory, converts
//     them (one MUFU.EX2 each) and parks them in a small per-lane staging area [frame][entry][lane].  The frame body
//     reads its emissions at [base + constant]: no address arithmetic, no bank conflicts.  (A first version gathered
//     the columns straight from global memory with 4-byte cp.async: 288 separate memory transactions per frame, and the
//     load / store unit, not HBM, was the bound: C5 forward 0.45 ms.)
//   * the gradient row is written in two passes like the log-domain kernels do: dense exp(x) * g for all V columns
//     (128-bit shared-memory loads of the row that is still in the ring, 128-bit stores when the gradient is
//     aligned), then the <= L + 1 columns that carry posterior mass.  Posterior mass is accumulated per USED column
//     (compact ids from a one-off bin-and-scan over the vocabulary), in the same fixed point with shared-memory
//     integer atomics.  In backward() a ring slot is refilled with a row of the NEXT tile as soon as its frame is done,
//     so the loads of a tile overlap the live phase of the one before.
#pragma once

namespace ssak {
namespace lin32 {

constexpr int LV_MAXV = 2048;       // (the log-domain kernels, which take the handed-back utterances, stop at ~2040)
constexpr int LV_MAXK = 7;          // shared-memory budget: row ring + tile + staging per warp
constexpr int LV_FROWS = 2 * C, LV_BROWS = C, LV_MC = 2;

struct LvSmem {
    int bars, ring, tile, stage, ucol, mass, total;
};
__host__ __device__ inline LvSmem lv_smem_map(int K, int V, bool grad) {
    LvSmem m;
    int o = 0;
    m.bars = o;  o += 16;                                                  // two mbarriers
    m.ring = o;  o += (grad ? LV_BROWS : LV_FROWS) * V * 4;                // whole rows (also the setup scratch: V ints)
    m.tile = o;  if (grad) o += C * 2 * K * 32 * 4;
    m.stage = o; o += C * (K + 3) * 32 * 4;                                // my entries of one chunk / tile
    m.ucol = o;  if (grad) o += ((32 * K + 2) * 4 + 15) & ~15;
    m.mass = o;  if (grad) o += (LV_MC * (32 * K + 2) * 4 + 15) & ~15;
    m.total = (o + 127) & ~127;
    return m;
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// my entries of ring row `src` (V raw log-probabilities / logits) -> emissions in the staging row d (+ lane):
// entry e <= K: label q0 - 1 + e, K + 1: blank, K + 2: the row normaliser (kept raw).  Returns the largest scaled
// emission seen (log2 units; > 0: not a probability).
template <int K>
__device__ __forceinline__ float lv_extract(float *d, const float *src, float z, const int (&lcol)[K + 1], int blank) {
    float xmax = -1.f;
#pragma unroll
    for (int e = 0; e <= K + 1; ++e) {
        const bool valid = e == K + 1 || lcol[e] >= 0;
        const int col = e == K + 1 ? blank : (lcol[e] >= 0 ? lcol[e] : 0);
        const float xs = fmaf(src[col], kLog2e, z);
        if (valid) xmax = fmaxf(xmax, xs);
        d[e * 32] = valid ? ex2_approx(xs) : 0.f;
    }
    d[(K + 2) * 32] = z;
    return xmax;
}

// vocabulary column of label q0 - 1 + e (e = 0 .. K), -1 where that label does not exist
template <int K>
__device__ __forceinline__ void make_columns(int q0, int L, int V, const int32_t *tg, int (&lcol)[K + 1]) {
#pragma unroll
    for (int e = 0; e <= K; ++e) {
        const int li = q0 - 1 + e;
        lcol[e] = -1;
        if (li >= 0 && li < L) {
            const int l = tg[li];
            lcol[e] = l < 0 ? 0 : (l >= V ? V - 1 : l);   // (memory safety; the join kernel turns the likelihood into NaN)
        }
    }
}

// the frame of step() with the lane's own gathered emissions: y = staging row + lane, entry e at y[32 e]
template <int K, int D, int MODE>
__device__ __forceinline__ void step_lv(float (&a)[K], float (&l)[K], const float ebp, const float *y, const float (&sk)[K],
                                        const float cin, float *tl, unsigned *mass, const int (&cidx)[K + 1], float &sbl) {
    float carry = cin;
#pragma unroll
    for (int kk = 0; kk < K; ++kk) {
        const int k = D ? K - 1 - kk : kk;
        const float el = y[(k + 1 - D) * 32];
        const float A = fmaf(a[k], ebp, carry);
        const float t = fmaf(sk[k], carry, fmaf(a[k], ebp, l[k]));
        if (MODE == 1) {
            tl[k * 32] = a[k];
            tl[(K + k) * 32] = carry;
        }
        if (MODE == 2) {
            sbl = fmaf(A, tl[k * 32], sbl);
            atomicAdd(reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(mass) + cidx[k + 1 - D]),
                      __float2uint_rn(t * tl[(K + k) * 32]));
        }
        carry = l[k];
        a[k] = A;
        l[k] = t * el;
    }
}

// ------------------------------------------------------------------------------------------------ forward
// ring of 2 x C rows: chunk n lives in slots (n & 1) * C + i, one mbarrier per half
template <int K>
__global__ void __launch_bounds__(32, FWD_WARPS) lv_forward_kernel(const Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    Chain ch;
    if (!chain_of(p, ch)) return;
    const int b = ch.b, dir = ch.dir, Tb = ch.Tb, L = ch.L;
    const int V = p.V;
    const int nrows = dir ? Tb - ch.m : ch.m;
    const float *lp_b = p.lp + (int64_t)b * p.sb;
    const float *zl_b = p.zl ? p.zl + b : nullptr;
    const int32_t *tg = p.targets + p.tgt_off[b];
    const LvSmem sm = lv_smem_map(K, V, false);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + sm.bars);
    float *ring = reinterpret_cast<float *>(smem + sm.ring);
    float *stage = reinterpret_cast<float *>(smem + sm.stage) + lane;   // entry (i, e) at (i * (K + 3) + e) * 32
    const int q0 = lane * K;
    const bool live = q0 <= L;
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto run = [&](auto dtag) {
        constexpr int D = decltype(dtag)::value;
        int lcol[K + 1], nocid[K + 1];
        float sk[K], as[K], ls[K];
        make_columns<K>(q0, L, V, tg, lcol);
        make_skip<K>(q0, L, V, tg, sk);
        start_row<K>(D, q0, L, as, ls);
        int E = 0;
        float f = 0.f, xmax = -1.f, dummy = 0.f, ebp = 1.f;
        bool ok = rescale<K, D>(as, ls, E, f, lane, EZERO, true);
        float *ck_dir = p.ck + ((int64_t)b * 2 + D) * p.NCK * p.ck_row;
        const int n_seq = (nrows + C - 1) / C;
        const int tf = D ? Tb - 1 : 0, dt = D ? -1 : 1;      // frame of row 0, direction of time
        auto issue = [&](int n) {                            // (lane 0) the rows of chunk n -> ring half n & 1
            if (n >= n_seq) return;
            const int nr = nrows - n * C < C ? nrows - n * C : C;
            mbar_arrive_expect_tx(&bars[n & 1], (uint32_t)(nr * V * 4));
            for (int i = 0; i < nr; ++i)
                bulk_g2s(ring + (size_t)((n & 1) * C + i) * V, lp_b + (int64_t)(tf + dt * (n * C + i)) * p.st, (uint32_t)(V * 4),
                         &bars[n & 1]);
        };
        if (lane == 0) {
            issue(0);
            issue(1);
        }
        for (int n = 0; n < n_seq; ++n) {
            const int row0 = n * C, nr = nrows - row0 < C ? nrows - row0 : C;
            float zr[C];
#pragma unroll
            for (int i = 0; i < C; ++i)
                zr[i] = (zl_b != nullptr && i < nr) ? __ldg(zl_b + (int64_t)(tf + dt * (row0 + i)) * p.B) : 0.f;
            mbar_wait(&bars[n & 1], (uint32_t)((n >> 1) & 1));
#pragma unroll
            for (int i = 0; i < C; ++i)
                if (i < nr)
                    xmax = fmaxf(xmax, lv_extract<K>(stage + i * (K + 3) * 32, ring + (size_t)((n & 1) * C + i) * V, zr[i], lcol, p.blank));
            __syncwarp();                                    // every lane has taken its columns: the half can be refilled
            if (lane == 0) {
                fence_proxy_async();
                issue(n + 2);
            }
#pragma unroll 1
            for (int i = 0; i < nr; ++i) {
                const float *y = stage + i * (K + 3) * 32;
                const float eb = y[(K + 1) * 32];
                step_lv<K, D, 0>(as, ls, ebp, y, sk, carry_in<K, D>(ls, f), nullptr, nullptr, nocid, dummy);
                ebp = eb;
            }
            if (nr == C) {
                unsigned h;
                ok = rescale<K, D>(as, ls, E, f, lane, EZERO, false, &h) && ok;
                if (p.save) store_checkpoint<K>(ck_dir + (int64_t)(n + 1) * p.ck_row, as, ls, ebp, E, h, lane, live);
            }
        }
        store_row<K>(p.fr + ((int64_t)b * 2 + D) * p.ck_row, as, ls, ebp, E, lane, live);
        unsigned h = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) h = max(h, max(__float_as_uint(as[k]), __float_as_uint(ls[k])));
        ok = ok && h < 0x7f800000u && !(xmax > 0.01f);
        if (__any_sync(FULL, !ok) && lane == 0) atomicOr(&p.flags[b], 1);   // inf / NaN / emissions above 1
    };
    if (dir) run(std::integral_constant<int, 1>{}); else run(std::integral_constant<int, 0>{});
}

// ------------------------------------------------------------------------------------------------ backward
// ring of C rows: slot i <-> row i of the current tile; a slot is refilled with row i of the NEXT tile as soon as the
// live direction is done with its frame; one mbarrier per tile parity
template <int K>
__global__ void __launch_bounds__(32, BWD_WARPS) lv_backward_kernel(const Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    Chain ch;
    if (!chain_of(p, ch)) return;
    const int b = ch.b, dir = ch.dir, Tb = ch.Tb, L = ch.L;
    const int V = p.V;
    if (p.flags[b] & 5) return;                                // the log-domain kernels own this utterance (or nobody)
    const float nll = p.nll[b];
    const float gs = p.grad_out[b];
    // 128-bit stores of the dense gradient pass: every gradient row 16-byte aligned
    const bool vecg = ((p.gst | p.gsb) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.grad) & 15) == 0;
    auto fill_row = [&](float *g, float v) {
        if (vecg) {
            float4 *g4 = reinterpret_cast<float4 *>(g);
            for (int it = lane; it < (V >> 2); it += 32) g4[it] = make_float4(v, v, v, v);
        } else {
            for (int cc = lane; cc < V; cc += 32) g[cc] = v;
        }
    };
    {
        // trivial outcomes (as in ctc_lattice_kernel)
        const bool infeasible = nll == __int_as_float(0x7f800000);
        const bool isnan_ = nll != nll;
        if (infeasible || isnan_ || Tb == 0) {
            if (dir == 0) {
                const float fillv = (isnan_ || (infeasible && !p.zero_inf)) ? __int_as_float(0x7fc00000) : 0.f;
                for (int t = 0; t < (int)p.T; ++t) fill_row(p.grad + (int64_t)t * p.gst + (int64_t)b * p.gsb, t < Tb ? fillv : 0.f);
            }
            return;
        }
    }
    const int rdir = 1 - dir;
    const int nrows = rdir ? Tb - ch.m : ch.m;                 // rows of DR = frames this chain handles
    const float *lp_b = p.lp + (int64_t)b * p.sb;
    const float *zl_b = p.zl ? p.zl + b : nullptr;
    const int32_t *tg = p.targets + p.tgt_off[b];
    const LvSmem sm = lv_smem_map(K, V, true);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + sm.bars);
    float *ring = reinterpret_cast<float *>(smem + sm.ring);       // [C][V]
    float *tile = reinterpret_cast<float *>(smem + sm.tile);       // [C][2K][32]
    float *stage = reinterpret_cast<float *>(smem + sm.stage) + lane;
    int *ucol = reinterpret_cast<int *>(smem + sm.ucol);           // [n_used] vocabulary column of compact id u, ascending
    unsigned *mass = reinterpret_cast<unsigned *>(smem + sm.mass); // [LV_MC][MS] posterior mass per used column, fixed point
    constexpr int MS = 32 * K + 2;
    unsigned *mass_mine = mass + (lane / (32 / LV_MC)) * MS;
    const int q0 = lane * K;
    const bool live = q0 <= L;
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }

    auto run = [&](auto dtag) {
        constexpr int DL = decltype(dtag)::value, DR = 1 - DL;
        int lcol[K + 1], cidx[K + 1], nocid[K + 1];
        float sk[K], la[K], ll[K], ra[K], rl[K];
        make_columns<K>(q0, L, V, tg, lcol);
        make_skip<K>(q0, L, V, tg, sk);
        // ---- compact ids of the used columns (labels of this utterance + blank), ascending: bin, scan (scratch: the
        //      row ring, V ints)
        int n_used, blank_cid;
        {
            int *bins = reinterpret_cast<int *>(ring);
            for (int cc = lane; cc < V; cc += 32) bins[cc] = 0;
            __syncwarp();
            for (int i = lane; i < L; i += 32) {
                int l = tg[i];
                l = l < 0 ? 0 : (l >= V ? V - 1 : l);
                bins[l] = 1;
            }
            if (lane == 0) bins[p.blank] = 1;
            __syncwarp();
            int run_ = 0;
            for (int c0 = 0; c0 < V; c0 += 32) {
                const int cc = c0 + lane;
                const int fl = cc < V ? bins[cc] : 0;
                const unsigned bal = __ballot_sync(FULL, fl != 0);
                const int id = run_ + __popc(bal & ((1u << lane) - 1u));
                if (fl) {
                    ucol[id] = cc;
                    bins[cc] = id;
                }
                run_ += __popc(bal);
            }
            n_used = run_;
            __syncwarp();
#pragma unroll
            for (int e = 0; e <= K; ++e) cidx[e] = 4 * (lcol[e] >= 0 ? bins[lcol[e]] : n_used);   // (n_used: the dump slot)
            blank_cid = bins[p.blank];
            for (int u = lane; u < LV_MC * MS; u += 32) mass[u] = 0u;
            __syncwarp();                                       // bins (the ring) are free now
        }
        int EL = 0, ER = 0;
        float fL = 0.f, fR = 0.f, dummy = 0.f, ebpL = 1.f;
        load_row<K>(p.fr + ((int64_t)b * 2 + DL) * p.ck_row, la, ll, EL, lane, live);
        const double log2P = -p.nll2[b];
        const double Epd = floor(log2P);
        const float invPm = (float)exp2(Epd - log2P);           // 1 / mantissa of P, in (0.5, 1]
        const int Ep = (int)Epd;
        const float *ck_dir = p.ck + ((int64_t)b * 2 + DR) * p.NCK * p.ck_row;
        const int n_seq = nrows > 0 ? nrows / C + 1 : 0;
        auto geometry = [&](int n, int &j, int &row0, int &i0, int &nr) {
            j = nrows / C - n;
            row0 = j * C - 1;
            i0 = j == 0 ? 1 : 0;
            nr = nrows - row0 < C ? nrows - row0 : C;
        };
        auto frame_of = [&](int rho) { return DR ? Tb - 1 - rho : rho; };
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
        // (lane 0) announce the rows of tile n on its mbarrier / fetch its row i into slot i
        auto expect_tile = [&](int n) {
            int j, row0, i0, nr;
            geometry(n, j, row0, i0, nr);
            mbar_arrive_expect_tx(&bars[n & 1], (uint32_t)((nr - i0) * V * 4));
        };
        auto fetch_row = [&](int n, int i) {
            int j, row0, i0, nr;
            geometry(n, j, row0, i0, nr);
            if (i >= i0 && i < nr)
                bulk_g2s(ring + (size_t)i * V, lp_b + (int64_t)frame_of(row0 + i) * p.st, (uint32_t)(V * 4), &bars[n & 1]);
        };
        bool bad = false;
        float xmax = -1.f;
        if (n_seq > 0) {
            if (lane == 0) {
                fence_proxy_async();                            // (the ring was written as scratch)
                expect_tile(0);
                for (int i = 0; i < C; ++i) fetch_row(0, i);
            }
            load_ck(nrows / C);
        }
        for (int n = 0; n < n_seq; ++n) {
            int j, row0, i0, nr;
            geometry(n, j, row0, i0, nr);
            float zr[C];
#pragma unroll
            for (int i = 0; i < C; ++i)
                zr[i] = (zl_b != nullptr && i >= i0 && i < nr) ? __ldg(zl_b + (int64_t)frame_of(row0 + i) * p.B) : 0.f;
            mbar_wait(&bars[n & 1], (uint32_t)((n >> 1) & 1));
#pragma unroll
            for (int i = 0; i < C; ++i)
                if (i >= i0 && i < nr)
                    xmax = fmaxf(xmax, lv_extract<K>(stage + i * (K + 3) * 32, ring + (size_t)i * V, zr[i], lcol, p.blank));
            if (j - 4 >= 1 && lane * 32 < p.ck_row)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ck_dir + (int64_t)(j - 4) * p.ck_row + lane * 32));
            __syncwarp();                                       // (also: the previous tile's gradient pass is done with the tile)
            // the next tile's rows: announced now; slots this tile does not use are fetched right away, the others
            // as their frames finish
            const bool more = n + 1 < n_seq;
            if (more && lane == 0) {
                fence_proxy_async();
                expect_tile(n + 1);
                for (int i = 0; i < C; ++i)
                    if (i < i0 || i >= nr) fetch_row(n + 1, i);
            }
            drop_dead<K>(ra, rl, ER);                           // (first use of the checkpoint fetched a tile ago)
            {
                const int Emin = ER > EZERO / 2 ? GMIN - ER + Ep : EZERO;
                bad = !rescale<K, DL>(la, ll, EL, fL, lane, Emin, n == 0) || bad;
            }
            {
                int g = EL + ER - Ep + FIX;
                g = g > GMAX ? GMAX : (g < -127 ? -127 : g);
                const float gfac = pow2f(g) * invPm;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    ra[k] *= gfac;
                    rl[k] *= gfac;
                }
                const int xr = ER - g;
                const int xu = DR ? __shfl_down_sync(FULL, xr, 1) : __shfl_up_sync(FULL, xr, 1);
                fR = lane == (DR ? 31 : 0) ? 0.f : pow2f(xu - xr);
            }
            // ---- R phase (see lin32_backward_kernel)
            float *tl = tile + lane;
            float ebpR = 1.f;
#pragma unroll 1
            for (int i = 0; i + 1 < nr; ++i) {
                const float *y = stage + (i + 1) * (K + 3) * 32;
                const float eb = y[(K + 1) * 32];
                step_lv<K, DR, 1>(ra, rl, ebpR, y, sk, carry_in<K, DR>(rl, fR), tl + i * 2 * K * 32, nullptr, nocid, dummy);
                ebpR = eb;
            }
            {
                const float cin = carry_in<K, DR>(rl, fR);
                float *ti = tl + (nr - 1) * 2 * K * 32;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    ti[k * 32] = ra[k];
                    const float c = DR ? (k == K - 1 ? cin : rl[k + 1 < K ? k + 1 : k]) : (k == 0 ? cin : rl[k > 0 ? k - 1 : 0]);
                    ti[(K + k) * 32] = c;
                }
            }
            if (more) load_ck(j - 1);
            // ---- B phase: live direction, posteriors, gradient row (dense pass from the ring, then the used columns)
#pragma unroll 1
            for (int i = nr - 1; i >= i0; --i) {
                const float *y = stage + i * (K + 3) * 32;
                const float eb = y[(K + 1) * 32];
                const float z = y[(K + 2) * 32];
                float sbl = 0.f;
                step_lv<K, DL, 2>(la, ll, ebpL, y, sk, carry_in<K, DL>(ll, fL), tl + i * 2 * K * 32, mass_mine, cidx, sbl);
                ebpL = eb;
                sbl *= i == 0 ? 1.f : eb;                       // (slot 0: the checkpoint's true blank states)
                bad = bad || !(sbl <= 3.0e38f);                 // inf / NaN
                unsigned tot = __float2uint_rn(sbl);
                const unsigned blank_mass = __reduce_add_sync(FULL, tot);
                __syncwarp();                                   // the frame's label posteriors are in mass[]
                const float *src = ring + (size_t)i * V;
                float *grow = p.grad + (int64_t)frame_of(row0 + i) * p.gst + (int64_t)b * p.gsb;
                // dense pass first (its stores are in flight while the masses are collected)
                if (vecg) {
                    const float4 *row4 = reinterpret_cast<const float4 *>(src);
                    float4 *g4 = reinterpret_cast<float4 *>(grow);
                    for (int it = lane; it < (V >> 2); it += 32) {
                        const float4 x = row4[it];
                        g4[it] = make_float4(ex2_approx(fmaf(x.x, kLog2e, z)) * gs, ex2_approx(fmaf(x.y, kLog2e, z)) * gs,
                                             ex2_approx(fmaf(x.z, kLog2e, z)) * gs, ex2_approx(fmaf(x.w, kLog2e, z)) * gs);
                    }
                } else {
                    for (int cc = lane; cc < V; cc += 32) grow[cc] = ex2_approx(fmaf(src[cc], kLog2e, z)) * gs;
                }
                unsigned mi[K + 1];
                tot = 0u;
#pragma unroll
                for (int jj = 0; jj <= K; ++jj) {
                    const int u = lane + 32 * jj;
                    mi[jj] = 0u;
                    if (u < n_used) {
                        unsigned msum = u == blank_cid ? blank_mass : 0u;
#pragma unroll
                        for (int c2 = 0; c2 < LV_MC; ++c2) {
                            msum += mass[c2 * MS + u];
                            mass[c2 * MS + u] = 0u;
                        }
                        mi[jj] = msum;
                    }
                    tot += mi[jj];
                }
                const float total = (float)__reduce_add_sync(FULL, tot);      // ~2^FIX
                bad = bad || !(fabsf(total * (1.0f / (float)(1 << FIX)) - 1.0f) <= p.mass_tol);
                const float inv = __fdividef(1.0f, total);
                __syncwarp();                                   // orders the dense stores before the overwrites below
#pragma unroll
                for (int jj = 0; jj <= K; ++jj) {
                    const int u = lane + 32 * jj;
                    if (u < n_used) {
                        const int cc = ucol[u];
                        grow[cc] = (ex2_approx(fmaf(src[cc], kLog2e, z)) - (float)mi[jj] * inv) * gs;
                    }
                }
                __syncwarp();                                   // mass[] is clean, and everybody is done with ring slot i
                if (more && lane == 0) {
                    fence_proxy_async();
                    fetch_row(n + 1, i);
                }
            }
        }
        bad = bad || xmax > 0.01f;
        if (__any_sync(FULL, bad) && lane == 0) {
            const int old = atomicOr(&p.flags[b], 2);
            if (!(old & 2)) {
                const int sl = atomicAdd(p.slot_counter, 1);
                if (sl < p.n_slots) { p.slot[b] = sl; p.slot_b[sl] = b; } else atomicOr(&p.flags[b], 4);
            }
        }
    };
    if (dir) run(std::integral_constant<int, 1>{}); else run(std::integral_constant<int, 0>{});
    if (dir == 0)   // frames beyond the utterance: exact zeros
        for (int t = Tb; t < (int)p.T; ++t) fill_row(p.grad + (int64_t)t * p.gst + (int64_t)b * p.gsb, 0.f);
}

}  // namespace lin32
}  // namespace ssak
