"""Randomised parity sweep of every kernel family on small shapes: python tools/fuzz_gpu.py [seconds] [seed]

Loss: the library's knobs force the throughput kernels (SSAK_CTC_LIN32=1) or one of the log-domain variants
(wavefront / barrier forward, few / many CTAs, posterior warps on / off) on shapes the default dispatch would not give
them -- T <= 260, targets up to 415 labels (also longer than the input), V from 2 to 1024, log-probabilities or raw
logits, every reduction; occasionally T <= 900 with targets up to 1300 labels -- against torch's CPU kernel in fp64.
Bars: loss 1e-5 relative (absolute below 1) + one fp32 rounding per frame, gradient 1e-4; both widened to 1.5 x the
error of torch's own fp32 CPU kernel on the same batch, the gradient also to 1e-6 x the largest likelihood (nats),
where those are larger.  Aligner: lane / wavefront / barrier kernels, planted / random / exact-tie emissions, with and
without first_as_garbage, bit-exact against the C oracle.  Greedy: argmax with exact ties + collapse.
A failing loss case is replayed with tools/fuzz_case.py / tools/fuzz_diag.py, an aligner case with
tools/fuzz_align_case.py; the cases found so far are pinned in tests/test_gpu_tight.py."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
import ssak_b200
from oracle import oracle as O
from ssak_b200.synth import align_batch, ctc_batch

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t_end, n_loss, n_align, n_greedy, worst, above, failures = time.time() + budget, 0, 0, 0, 0.0, [], []
KNOBS = ("SSAK_CTC_LIN32", "SSAK_CTC_FWD_WAVE", "SSAK_CTC_FEW", "SSAK_CTC_SPLIT", "SSAK_ALIGN_LANE", "SSAK_ALIGN_WAVE")


def set_knobs(**kw):
    """Kernel-family knobs of the library (read per call): None = the library's own choice."""
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in kw.items():
        if v is not None:
            os.environ[k] = str(v)
    return " ".join(f"{k}={v}" for k, v in kw.items() if v is not None)


while time.time() < t_end:
    # ---- loss: throughput kernels forced (half of the cases), or the log-domain kernels in each of their variants
    lin = int(rng.integers(0, 2))
    knobs = set_knobs(SSAK_CTC_LIN32=lin,
                      SSAK_CTC_FWD_WAVE=None if lin or rng.integers(0, 2) else 0,
                      SSAK_CTC_FEW=None if lin or rng.integers(0, 2) else 0,
                      SSAK_CTC_SPLIT=None if lin or rng.integers(0, 2) else 0)
    V = int(rng.choice([2, 5, 33, 50, 64, 65, 128, 132, 256, 512, 1024]))
    big = (not lin) and rng.integers(0, 6) == 0
    Lmax = int(rng.integers(0, 224 if V > 128 else 416)) if not big else int(rng.integers(300, 1300))
    T = int(rng.integers(1, 260)) if not big else int(rng.integers(300, 900))
    B = int(rng.integers(1, 9)) if not big else int(rng.integers(1, 4))
    if lin and rng.integers(0, 8) == 0:      # many chains at once: launch order, more hand-backs than a small batch
        B, T = int(rng.integers(20, 33)), min(T, 120)   # (<= 32: a row block for every utterance, see lin_slots)
    planted = bool(rng.integers(0, 2))
    seed = int(rng.integers(1 << 30))
    lp, tg, il, tl = ctc_batch(B, T, V, 0, Lmax, seed, Tmin=1, planted=planted and V > 2)
    tl = torch.minimum(tl, torch.tensor(Lmax))
    il = torch.clamp(il, 1, T)
    logits = bool(rng.integers(0, 3) == 0)
    red = ["none", "mean", "sum"][int(rng.integers(0, 3))]
    x0 = lp * 1.7 + 0.3 if logits else lp
    x = x0.cuda().requires_grad_(True)
    fn = ssak_b200.ctc_loss_from_logits if logits else ssak_b200.ctc_loss
    loss = fn(x, tg, il, tl, 0, red, True)
    loss.sum().backward()
    y = x0.double().requires_grad_(True)
    ref = F.ctc_loss(F.log_softmax(y, -1) if logits else y, tg, il, tl, 0, red, True)
    ref.sum().backward()
    # the reference itself (torch's fp32 CPU kernel) against the fp64 truth: the bar scales with its error where that
    # exceeds the 1e-4 / 1e-5 bars (likelihoods of ~1000 nats)
    y32 = x0.clone().requires_grad_(True)
    ref32 = F.ctc_loss(F.log_softmax(y32, -1) if logits else y32, tg, il, tl, 0, red, True)
    ref32.sum().backward()
    ref_gerr = (y32.grad.double() - y.grad).abs().max().item()
    ref_lerr = ((ref32.detach().double() - ref.detach()).abs() / ref.detach().abs().clamp_min(1.0)).reshape(-1)
    ref_lerr = ref_lerr[torch.isfinite(ref_lerr)].max().item() if torch.isfinite(ref_lerr).any() else 0.0
    l, r = loss.detach().cpu().double().reshape(-1), ref.detach().reshape(-1)
    nll_max = float(F.ctc_loss(F.log_softmax(y.detach(), -1) if logits else y.detach(), tg, il, tl, 0, 'none', True).max())
    case = f"{knobs} python tools/fuzz_case.py {V} {Lmax} {T} {B} {int(planted)} {int(logits)} {seed}   # reduction {red}"
    if not torch.equal(torch.isfinite(l), torch.isfinite(r)):
        failures.append(("finite", case)); continue
    fin = torch.isfinite(r)
    if fin.any():
        # 1e-5 relative (absolute below 1), or 1.5 x the reference's own error, + one fp32 rounding of an O(1) term per
        # frame (the row normalisers of the logits path, the emissions)
        frames = float(il.max()) if red == "none" else float(il.sum())
        tol = max(1e-5, 1.5 * ref_lerr + 1e-5) * r[fin].abs().clamp_min(1.0) + 2e-7 * frames
        bad = ((l - r).abs()[fin] > tol)
        if bad.any():
            rel = ((l - r).abs() / r.abs().clamp_min(1.0))[fin].max().item()
            failures.append((f"loss {rel:.2e}", case)); continue
    err = (x.grad.cpu().double() - y.grad).abs().max().item()
    if err > max(1e-4, 1.5 * ref_gerr + 1e-5, 1e-6 * nll_max):
        failures.append((f"grad {err:.2e}", case)); continue
    if err > 1e-5:
        above.append((round(err, 7), V, Lmax, T, B, planted, logits))
    worst = max(worst, err)
    n_loss += 1
    # ---- aligner: the throughput (one warp per utterance) kernel, the wavefront kernel, the barrier kernel
    ak = int(rng.integers(0, 3))
    aknobs = set_knobs(SSAK_ALIGN_LANE=1 if ak == 0 else 0, SSAK_ALIGN_WAVE=0 if ak == 2 else None)
    V = int(rng.choice([3, 7, 50, 64, 97, 128]))
    Lmax = int(rng.integers(1, 512))
    T = int(rng.integers(1, 700))
    B = int(rng.integers(1, 7))
    kind = ["planted", "random", "tie"][int(rng.integers(0, 3))]
    fag = bool(rng.integers(0, 2))
    aseed = int(rng.integers(1 << 30))
    em, toks, el, tl2 = align_batch(B, T, V, 1, Lmax, aseed, Tmin=1, kind=kind)
    res = ssak_b200.forced_align(em.cuda(), toks, el, tl2, first_as_garbage=fag)
    st, en, ts, status = res.starts.cpu(), res.ends.cpu(), res.t_start.cpu(), res.status.cpu()
    for b in range(B):
        Tb, Lb = int(el[b]), int(tl2[b])
        if fag and Lb > 0 and np.isnan(O.garbage_col0(em[b, :Tb].numpy(), int(toks[b, 0]))).any():
            continue   # an emission above 0 (the un-normalised "tie" rows): log(1 - exp(e)) is NaN, the result undefined
        rc, ss, se, sc, t0 = O.align(em[b, :Tb].numpy(), toks[b, :Lb].tolist(), 0, fag)
        good = (rc == 0) == (int(status[b]) == 0)
        if good and rc == 0:
            good = int(ts[b]) == t0 and st[b, :Lb].tolist() == ss.tolist() and en[b, :Lb].tolist() == se.tolist()
        if not good:
            failures.append(("align", f"{aknobs} V={V} Lmax={Lmax} T={T} B={B} kind={kind} first_as_garbage={fag} seed={aseed} b={b}"))
            break
    n_align += 1
    # ---- greedy: argmax (first maximum) + collapse, rows with exact ties
    if n_align % 4 == 0:
        import itertools
        V, T, B = int(rng.choice([1, 3, 33, 50, 257, 1024])), int(rng.integers(1, 90)), int(rng.integers(1, 6))
        gseed = int(rng.integers(1 << 30))
        gg = torch.Generator().manual_seed(gseed)
        pr = torch.round(torch.randn(B, T, V, generator=gg) * 2) / 2
        nfr = torch.randint(0, T + 1, (B,), generator=gg)
        blank = int(rng.integers(0, V))
        ids, out, lens = ssak_b200.greedy_ids(pr.cuda(), nfr, blank)
        good = torch.equal(ids.cpu().long(), torch.argmax(pr, -1))
        for b in range(B):
            exp = [k for k, _ in itertools.groupby(torch.argmax(pr[b, : int(nfr[b])], -1).tolist()) if k != blank]
            good = good and out[b, : int(lens[b])].cpu().tolist() == exp
        if not good:
            failures.append(("greedy", f"V={V} T={T} B={B} blank={blank} seed={gseed}"))
        n_greedy += 1
for f in failures:
    print("FAIL", f)
print(f"fuzz {'ok' if not failures else 'FAILED'}: {n_loss} loss batches (worst gradient error {worst:.1e}; above 1e-5 -- utterances recomputed by the "
      f"log-domain kernels: {above[:8]}), {n_align} aligner batches, {n_greedy} greedy batches")
