"""Numerics prototype (numpy, CPU) of the block-floating-point linear-domain CTC recursion used by
ssak_b200/csrc/ctc_lin32.cu: fp32 mantissas, one integer exponent per LANE (K consecutive positions), flush to
zero below 2^-126, re-scaling every C frames with a decaying prefix-max scan of the lane exponents.

Checks, against an fp64 log-domain reference: log P, the posteriors (-> gradient error), and the per-frame mass
sum_s gamma_t(s) (the quantity the kernel uses to detect states lost to the fp32 range).

    python tools/proto_bfp.py [case ...]      cases: planted random flat deletion
"""
import sys

import numpy as np

F = np.float32
TINY = F(2.0 ** -126)


def ftz(x):
    x = x.astype(F, copy=False)
    x[np.abs(x) < TINY] = 0
    return x


def pow2(e):
    """2^e as fp32 with the exponent clamped like the kernel does (below -126 -> 0)."""
    e = np.asarray(e, dtype=np.int64)
    out = np.ldexp(np.ones(e.shape, dtype=np.float64), np.clip(e, -200, 127)).astype(F)
    out[e < -126] = 0
    return out


class Chain:
    """One direction of one utterance.  D = 0: alpha (position q = blank q, label q; flow towards higher lanes),
    D = 1: beta (position q = label q-1, blank q; flow towards lower lanes)."""

    def __init__(self, D, labels, V, K, C, T0, DMAX):
        self.D, self.K, self.C, self.T0, self.DMAX = D, K, C, T0, DMAX
        L = len(labels)
        self.L = L
        P = 32 * K
        assert L + 1 <= P
        q = np.arange(P)
        li = q - 1 if D else q                      # label index of the position's label state
        ok = (li >= 0) & (li < L) & (q <= L)
        self.lab = np.where(ok, np.asarray(labels + [0] * (P + 1))[np.clip(li, 0, L)], V)   # V = zero column
        sk = np.zeros(P, dtype=F)
        for qq in range(1, L):
            if labels[qq - 1] != labels[qq]:
                sk[qq] = 1
        self.sk = sk
        self.b = np.zeros(P, dtype=F)
        self.l = np.zeros(P, dtype=F)
        self.b[L if D else 0] = 1
        self.E = np.zeros(32, dtype=np.int64)
        self.f = np.ones(32, dtype=F)               # factor applied to the carry that enters the lane
        self.n = 0
        self.lane = q // K
        self.rescale()

    def carry(self):
        """label state of the neighbouring position (previous frame), in the receiving lane's scale"""
        K = self.K
        if self.D == 0:
            c = np.concatenate([[F(0)], self.l[:-1]])
            first = np.arange(0, 32 * K, K)
        else:
            c = np.concatenate([self.l[1:], [F(0)]])
            first = np.arange(K - 1, 32 * K, K)
        c = c.copy()
        c[first] = ftz(c[first] * self.f)
        return c

    def step(self, y, blank):
        """y: [V+1] fp32 emissions of the frame (last entry 0).  Returns (A, t, carry): the states before their
        emission is applied."""
        c = self.carry()
        A = ftz(self.b + c)
        t = ftz(ftz(self.l + self.b) + self.sk * c)
        self.b = ftz(A * y[blank])
        self.l = ftz(t * y[self.lab])
        self.n += 1
        if self.n % self.C == 0:
            self.rescale()
        return A, t, c

    def rescale(self):
        K = self.K
        m = np.maximum(self.b, self.l).reshape(32, K).max(1)
        with np.errstate(divide="ignore"):
            e_own = np.where(m > 0, np.floor(np.log2(np.maximum(m, TINY).astype(np.float64))), -10 ** 6).astype(np.int64)
        need = np.where(m > 0, self.E + e_own - self.T0, -10 ** 6)
        new = need.copy()
        rng = range(1, 32) if self.D == 0 else range(30, -1, -1)
        prev = (lambda j: j - 1) if self.D == 0 else (lambda j: j + 1)
        for j in rng:
            new[j] = max(new[j], new[prev(j)] - self.DMAX)
        d = self.E - new                                   # mantissas *= 2^d
        fac = pow2(d)
        self.b = ftz(self.b * np.repeat(fac, K))
        self.l = ftz(self.l * np.repeat(fac, K))
        self.E = new
        nb = np.roll(new, 1 if self.D == 0 else -1)
        self.f = pow2(nb - new)
        self.f[0 if self.D == 0 else 31] = 0


def reference(lp, labels, blank):
    """fp64 log-domain alpha / beta / posteriors."""
    T, V = lp.shape
    L = len(labels)
    S = 2 * L + 1
    ext = np.full(S, blank)
    ext[1::2] = labels
    NEG = -np.inf
    la = np.full((T, S), NEG)
    lb = np.full((T, S), NEG)
    skip = np.zeros(S, dtype=bool)
    skip[3::2] = ext[3::2] != ext[1:-2:2]
    la[0, 0] = lp[0, blank]
    if S > 1:
        la[0, 1] = lp[0, ext[1]]
    for t in range(1, T):
        p0 = la[t - 1]
        p1 = np.concatenate([[NEG], p0[:-1]])
        p2 = np.where(skip, np.concatenate([[NEG, NEG], p0[:-2]]), NEG)
        la[t] = np.logaddexp(np.logaddexp(p0, p1), p2) + lp[t, ext]
    lb[T - 1, S - 1] = lp[T - 1, blank]
    if S > 1:
        lb[T - 1, S - 2] = lp[T - 1, ext[S - 2]]
    skipb = np.zeros(S, dtype=bool)
    skipb[1:-2:2] = ext[1:-2:2] != ext[3::2]
    for t in range(T - 2, -1, -1):
        p0 = lb[t + 1]
        p1 = np.concatenate([p0[1:], [NEG]])
        p2 = np.where(skipb, np.concatenate([p0[2:], [NEG, NEG]]), NEG)
        lb[t] = np.logaddexp(np.logaddexp(p0, p1), p2) + lp[t, ext]
    logP = np.logaddexp(la[T - 1, S - 1], la[T - 1, S - 2] if S > 1 else NEG)
    gam = np.exp(la + lb - lp[:, ext] - logP)
    return logP, gam, ext


def run(lp, labels, blank=0, K=13, C=4, T0=20, DMAX=24):
    T, V = lp.shape
    L = len(labels)
    y = np.concatenate([np.exp2((lp.astype(F) * F(1.4426950408889634)).astype(F)), np.zeros((T, 1), F)], 1).astype(F)
    # alpha over all frames, keeping (A, t) rows with their exponents; beta likewise; posteriors = alpha_pre * beta_post
    a = Chain(0, labels, V, K, C, T0, DMAX)
    Apre = np.zeros((T, 32 * K)); Tpre = np.zeros((T, 32 * K)); Ea = np.zeros((T, 32), dtype=np.int64)
    for t in range(T):
        Ea[t] = a.E
        A, tt, _ = a.step(y[t], blank)
        Apre[t], Tpre[t] = A, tt
    # P from the final alpha row: blank L + label L-1 (post-emission)
    fin = (a.b[L].astype(np.float64) * 2.0 ** float(a.E[L // K]) +
           (a.l[L - 1].astype(np.float64) * 2.0 ** float(a.E[(L - 1) // K]) if L > 0 else 0.0))
    # (exponent may exceed double range for long utterances: use log2 arithmetic)
    def l2(m, e):
        return np.log2(m.astype(np.float64)) + e if m > 0 else -np.inf
    l2P = np.logaddexp2(l2(a.b[L], a.E[L // K]), l2(a.l[L - 1], a.E[(L - 1) // K]) if L > 0 else -np.inf)
    bch = Chain(1, labels, V, K, C, T0, DMAX)
    gam_b = np.zeros((T, L + 1)); gam_l = np.zeros((T, max(L, 1)))
    mass = np.zeros(T)
    lane = np.arange(32 * K) // K
    for t in range(T - 1, -1, -1):
        bch.step(y[t], blank)          # beta post-emission at frame t = state after the step (before any rescale? see below)
        # note: step() may have re-scaled; the post-emission state is (b, l) with exponent E either way
        eb = bch.E[lane] + Ea[t][lane]
        sc = np.exp2((eb - l2P).astype(np.float64))
        pb = Apre[t] * bch.b.astype(np.float64) * sc                    # blank q: alpha_pre(blank q) * beta(blank q)
        # label q: alpha position q (label q) pairs with beta position q+1 (label q) -- different lanes at the edges
        bl = np.concatenate([bch.l[1:], [0]]).astype(np.float64)
        el = np.concatenate([bch.E[lane][1:], [0]]) + Ea[t][lane]
        pl = Tpre[t] * bl * np.exp2((el - l2P).astype(np.float64))
        gam_b[t] = pb[: L + 1]
        if L > 0:
            gam_l[t] = pl[:L]
        mass[t] = pb[: L + 1].sum() + pl[:L].sum()
    return l2P * np.log(2.0), gam_b, gam_l, mass


def make_case(name, T=1500, L=300, V=50, seed=0):
    rng = np.random.default_rng(seed)
    labels = list(rng.integers(1, V, L))
    if name == "random":
        lg = rng.standard_normal((T, V))
    elif name == "flat":
        lg = 0.01 * rng.standard_normal((T, V))
    else:
        boost = 6.0 if name != "sharp" else 14.0
        lg = rng.standard_normal((T, V))
        lg[:, 0] += boost
        if name == "deletion":        # the audio lacks the labels [100, 100+ND): their onsets are missing
            ND = 25
            keep = [i for i in range(L) if not (100 <= i < 100 + ND)]
            onset = np.sort(rng.permutation(T)[: len(keep)])
            lg[onset, 0] -= boost
            lg[onset, np.asarray(labels)[keep]] += boost
        else:
            onset = np.sort(rng.permutation(T)[:L])
            lg[onset, 0] -= boost
            lg[onset, labels] += boost
    lp = lg - np.log(np.exp(lg).sum(1, keepdims=True))
    return lp.astype(F), [int(x) for x in labels]


if __name__ == "__main__":
    cases = sys.argv[1:] or ["planted", "random", "flat", "deletion", "sharp"]
    for name in cases:
        for (T, L) in ((1500, 300), (1500, 400), (700, 40)):
            lp, labels = make_case(name, T, L)
            logP, gam, ext = reference(lp.astype(np.float64), labels, 0)
            for (C, T0, DMAX) in ((4, 20, 24), (8, 20, 24)):
                lP, gb, gl, mass = run(lp, labels, K=(L + 1 + 31) // 32, C=C, T0=T0, DMAX=DMAX)
                err_b = np.abs(gb - gam[:, 0::2]).max()
                err_l = np.abs(gl - gam[:, 1::2]).max() if L else 0
                # per-frame normalised
                nb = np.abs(gb / mass[:, None] - gam[:, 0::2]).max()
                nl = np.abs(gl / mass[:, None] - gam[:, 1::2]).max()
                print(f"{name:9s} T={T} L={L} C={C} T0={T0} DMAX={DMAX}: logP rel {abs(lP - logP) / abs(logP):.2e} "
                      f"gamma err {max(err_b, err_l):.2e} normalised {max(nb, nl):.2e} "
                      f"mass dev max {np.abs(mass - 1).max():.2e}")
