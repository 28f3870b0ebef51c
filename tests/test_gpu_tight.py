"""Tight alignments (T_b ~ L_b + repeats) under peaky emissions that disagree with the transcript: the only feasible
paths run hundreds of bits below the row maximum.  Found by tools/fuzz_gpu.py; two defects are pinned here:
  * the wavefront forward of the log-domain kernels took all three terms of a pair over ONE maximum and flushed the
    blank sum / a repeated label's sum to log 0 when the third term dominated by > 2^126 (likelihood off by tens of
    nats, garbage gradient);
  * the throughput kernels' forward likelihood of such an utterance is only verified by the backward call's
    self-check: the likelihood is now final after backward (C ABI) / on return (ssak_b200.ctc_loss runs both calls).
Checked against torch CPU fp64; the gradient bar is the north star's 1e-4, or 1.5 x the error torch's own fp32 CPU
kernel (the reference) makes on the same utterance when that is larger, or 1e-6 x nll_b (the absolute rounding error
of an fp32 log-likelihood grows with its magnitude: 7e-4 at 1362 nats with no slack at all, torch fp32: 4e-4)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# V Lmax T B planted logits seed  (tools/fuzz_case.py replays one)
CASES = ["65 257 210 8 1 1 206911496", "33 228 152 8 1 1 269815783", "33 256 147 3 1 0 224398287",
         "33 121 122 8 1 1 971909417", "64 293 229 5 1 1 1071034586", "128 326 167 8 1 1 264862607",
         "64 301 259 3 1 1 1049753191", "50 205 186 3 1 0 29280320", "5 206 248 3 1 0 429124708",
         "132 181 200 8 0 1 581077147", "5 185 232 7 1 1 248227161", "50 381 93 8 1 1 179812443",
         "64 407 181 7 1 1 408795322", "50 349 195 3 1 0 508654026"]


def _inputs(case):
    from ssak_b200.synth import ctc_batch
    V, Lmax, T, B, planted, logits, seed = (int(v) for v in case.split())
    lp, tg, il, tl = ctc_batch(B, T, V, 0, Lmax, seed, Tmin=1, planted=bool(planted) and V > 2)
    tl = torch.minimum(tl, torch.tensor(Lmax))
    il = torch.clamp(il, 1, T)
    return (lp * 1.7 + 0.3 if logits else lp), tg, il, tl, bool(logits)


def _reference(x0, tg, il, tl, logits, dtype):
    y = x0.detach().clone().to(dtype).requires_grad_(True)
    ref = F.ctc_loss(F.log_softmax(y, -1) if logits else y, tg, il, tl, 0, "none", True)
    ref.sum().backward()
    return ref.detach().double(), y.grad.double()


def _compare(loss, grad, x0, tg, il, tl, logits):
    r64, g64 = _reference(x0, tg, il, tl, logits, torch.float64)
    r32, g32 = _reference(x0, tg, il, tl, logits, torch.float32)
    l = loss.detach().cpu().double()
    assert torch.equal(torch.isfinite(l), torch.isfinite(r64))
    # (+ one fp32 rounding of an O(1) term per frame: the logits path adds a row normaliser per frame)
    tol_l = torch.maximum(1e-5 * r64.abs().clamp_min(1.0), 1.5 * (r32 - r64).abs() + 1e-5) + 1e-7 * il.double()
    assert ((l - r64).abs() <= tol_l).all(), (l, r64)
    err = (grad.cpu().double() - g64).abs().amax(dim=(0, 2))
    ref_err = (g32 - g64).abs().amax(dim=(0, 2))
    bar = torch.maximum(torch.tensor(1e-4, dtype=torch.float64), 1.5 * ref_err + 1e-5)
    bar = torch.maximum(bar, 1e-6 * r64.abs())
    assert (err <= bar).all(), (err, ref_err)


@pytest.mark.parametrize("mode", ["0", "1"])
@pytest.mark.parametrize("case", CASES)
def test_tight_alignments(case, mode, monkeypatch):
    """mode 0: the log-domain kernels (the default for these batch sizes); 1: the throughput kernels forced."""
    import ssak_b200
    monkeypatch.setenv("SSAK_CTC_LIN32", mode)
    x0, tg, il, tl, logits = _inputs(case)
    x = x0.cuda().requires_grad_(True)
    fn = ssak_b200.ctc_loss_from_logits if logits else ssak_b200.ctc_loss
    loss = fn(x, tg, il, tl, 0, "none", True)
    loss.sum().backward()
    _compare(loss, x.grad, x0, tg, il, tl, logits)
    # forward only (no gradient wanted): always the log-domain kernels
    with torch.no_grad():
        loss2 = fn(x0.cuda(), tg, il, tl, 0, "none", True)
    r64, _ = _reference(x0, tg, il, tl, logits, torch.float64)
    assert ((loss2.cpu().double() - r64).abs() <= 2e-5 * r64.abs().clamp_min(1.0)).all()


def test_c_abi_likelihood_is_final_after_backward(monkeypatch):
    """Throughput kernels through the C ABI: an utterance whose forward likelihood lost states to the fp32 range is
    caught by backward's self-check (flag bit 1) and its neg_log_likelihood entry is rewritten."""
    from test_gpu_lin32 import _path_flags
    monkeypatch.setenv("SSAK_CTC_LIN32", "1")
    x0, tg, il, tl, logits = _inputs("50 349 195 3 1 0 508654026")
    assert not logits
    r64, g64 = _reference(x0, tg, il, tl, False, torch.float64)
    fl, nll, grad = _path_flags(x0, tg, il, tl)
    nll = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll).double()
    assert ((nll - r64).abs() <= 2e-5 * r64.abs().clamp_min(1.0)).all(), (nll, r64, fl)
    assert (fl & 4).sum() == 0


def test_eager_upstream_gradient(monkeypatch):
    """Eager mode (both calls inside forward): autograd's upstream gradient is applied afterwards -- a scalar for the
    reduced losses, a vector for 'none' -- and a second backward over a retained graph is consistent."""
    import ssak_b200
    from ssak_b200.synth import ctc_batch
    monkeypatch.setenv("SSAK_CTC_LIN32", "1")
    lp, tg, il, tl = ctc_batch(6, 120, 50, 5, 40, 77, Tmin=60)
    w = torch.tensor([0.5, 2.0, 1.0, -1.0, 3.0, 0.25])
    for red in ("none", "mean", "sum"):
        x = lp.cuda().requires_grad_(True)
        loss = ssak_b200.ctc_loss(x, tg, il, tl, 0, red, True)
        obj = (loss * w.cuda()).sum() if red == "none" else loss * 3.0
        obj.backward(retain_graph=True)
        g1 = x.grad.clone()
        x.grad = None
        obj.backward()
        y = lp.double().requires_grad_(True)
        ref = F.ctc_loss(y, tg, il, tl, 0, red, True)
        ((ref * w.double()).sum() if red == "none" else ref * 3.0).backward()
        assert (g1.cpu().double() - y.grad).abs().max() <= 2e-5
        assert (x.grad.cpu().double() - y.grad).abs().max() <= 2e-5


@pytest.mark.parametrize("reduction", ["mean", "sum", "mean_volume"])
def test_sharded_wrapper_eager(reduction, monkeypatch):
    """The sharded loss (world size 1) on the throughput kernels: both calls inside forward, the upstream gradient
    (and 'mean_volume's denominator) applied in backward; one utterance of the batch is a tight one whose
    provisional likelihood is revised."""
    from ssak_b200.shard import sharded_ctc_loss
    monkeypatch.setenv("SSAK_CTC_LIN32", "1")
    x0, tg, il, tl, logits = _inputs("50 349 195 3 1 0 508654026")
    x = x0.cuda().requires_grad_(True)
    loss = sharded_ctc_loss(x, tg.cuda(), il.cuda(), tl.cuda(), 0, reduction, True, global_batch=3)
    loss.backward(torch.tensor(0.5, device="cuda"))
    y = x0.double().requires_grad_(True)
    if reduction == "mean_volume":
        nll = F.ctc_loss(y, tg, il, tl, 0, "none", True)
        ref = nll.sum() / tl.sum()
    else:
        ref = F.ctc_loss(y, tg, il, tl, 0, reduction, True)
    (0.5 * ref).backward()
    assert abs(loss.item() - ref.item()) <= 2e-5 * max(abs(ref.item()), 1.0)
    assert (x.grad.cpu().double() - y.grad).abs().max().item() <= 1e-4


def test_default_dispatch_batch_with_tight_utterances():
    """B = 256 through the DEFAULT dispatch (throughput kernels from B = 222 on): 240 ordinary utterances and 16
    tight ones (170 .. 199 labels on 200 frames, planted peaks 1.7 x sharper) that end up on the log-domain kernels'
    row blocks -- some already in forward (likelihood 0 in fp32 block floating point), some in backward."""
    import ssak_b200
    from ssak_b200.synth import ctc_batch
    from test_gpu_lin32 import _path_flags
    T, V = 200, 33
    lp1, tg1, il1, tl1 = ctc_batch(240, T, V, 5, 60, 4242, Tmin=80)
    lp2, tg2, il2, tl2 = ctc_batch(16, T, V, 170, 199, 4243, Tmin=T)
    tg = torch.zeros(256, 199, dtype=tg1.dtype)
    tg[:240, :tg1.shape[1]] = tg1
    tg[240:, :tg2.shape[1]] = tg2
    x0 = torch.cat([lp1, (lp2 * 1.7).log_softmax(-1)], 1)
    il, tl = torch.cat([il1, il2]), torch.cat([tl1, tl2])
    perm = torch.randperm(256, generator=torch.Generator().manual_seed(5))
    x0, tg, il, tl = x0[:, perm].contiguous(), tg[perm], il[perm], tl[perm]
    assert ssak_b200.lib().ssak_ctc_loss_nll_is_provisional(256, V, int(tl.max())) == 1
    x = x0.cuda().requires_grad_(True)
    loss = ssak_b200.ctc_loss(x, tg, il, tl, 0, "none", True)
    loss.sum().backward()
    _compare(loss, x.grad, x0, tg, il, tl, False)
    fl, _, _ = _path_flags(x0, tg, il, tl)
    tight = (tl >= 170)
    assert (fl[~tight] == 0).all() and (fl & 4).sum() == 0
    print("path flags of the tight utterances:", fl[tight].tolist())
