// ctc_lin32.h -- interface between ctc_loss.cu (dispatch, workspace layout, log-domain kernels) and ctc_lin32.cu
// (the throughput kernels: linear-domain block-floating-point recursion, one warp per (utterance, direction)).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ssak {
namespace lin32 {

constexpr int C = 4;            // frames per chunk = checkpoint spacing = rows per backward tile
constexpr int MAXK = 13;        // positions per lane: targets up to 32 * 13 - 1 = 415 labels
constexpr int MAXV = 128;       // vocabulary columns a warp converts per frame (4 per lane); beyond: the gather kernels
                                // (ctc_lin32_lv.cuh: V <= 2048, targets up to 223 labels)

struct Params {
    const float *lp;            // log-probabilities, or raw logits when zl != nullptr
    int64_t T, B;
    int V;
    int64_t st, sb;
    const int32_t *targets;
    const int64_t *tgt_off;
    const int32_t *in_len;
    const int32_t *tgt_len;
    int Lmax, blank;
    const float *zl;            // [T][B] -log2 sum exp (logits entry points) or nullptr
    float *ck;                  // [B][2][NCK][ck_row] checkpoint j = state after j*C steps (j >= 1)
    float *fr;                  // [B][2][ck_row] frontier rows (state after all forward steps of the direction)
    int NCK, ck_row;            // ck_row = 64 K + 32: blank states | label states (each [K][32]) | lane exponents
    double *nll2;               // [B] -log2 P
    float *nll;                 // [B]
    int *flags;                 // [B] bit 0: handed to the log-domain kernels by forward(); bit 1: by backward();
                                //     bit 2: nobody owns it (no row block left): NaN gradient
    int *slot;                  // [B] row block of a handed-back utterance in the log-domain kernels' `rows`
    int *slot_b;                // [n_slots] the utterance that owns row block i (the masked log-domain launches run one
                                //     CTA pair per row block)
    int *slot_counter;
    int n_slots;
    int *order;                 // [B] utterances by decreasing T_b (chains 2r, 2r+1 belong to utterance order[r]: the
                                //     longest chains start first, the partial last wave holds the shortest), or nullptr
    const float *grad_out;
    float *grad;
    int64_t gst, gsb;
    int zero_inf;
    int save;
    int K;                      // positions per lane (instantiated: 4, 7, 10, 13)
    float mass_tol;             // |sum of the posteriors of a frame - 1| beyond which backward() hands the utterance back
};

// positions per lane for targets up to Lmax labels, 0 when the kernels do not cover the shape
int lanes_k(int64_t Lmax, int64_t V);
bool ordered(int64_t B);        // the launches take the utterances by decreasing T_b (Params::order is used)
inline int ck_row_elems(int K) { return 64 * K + 32; }
inline int n_checkpoints(int64_t T) { return (int)((T / 2 + 1) / C) + 2; }

// forward: recursion kernel + join kernel (nll, nll2, flags, slots).  backward: recursion + gradient kernel.
int launch_forward(const Params &p, cudaStream_t s);
int launch_backward(const Params &p, cudaStream_t s);
int launch_orphans(const Params &p, cudaStream_t s);   // NaN gradient for the utterances nobody owns (after the fallback launches)

}  // namespace lin32
}  // namespace ssak
