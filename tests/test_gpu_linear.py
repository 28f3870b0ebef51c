"""The linear-domain fp64 loss kernels (ssak_b200/csrc/ctc_lin.cuh, opt-in: SSAK_CTC_LINEAR=1): no stored lattice
(checkpoint every 8 frames + recomputation), same results as torch's CPU ctc_loss in fp64 -- and closer to it than any
fp32 log-domain recursion -- including the utterances the kernels hand back to the log-domain path (fp64 range)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-5
GRAD_ATOL = 2e-6          # the linear-domain path is exact to fp32 rounding of the emissions


@pytest.fixture(autouse=True)
def _linear(monkeypatch):
    monkeypatch.setenv("SSAK_CTC_LINEAR", "1")


def _check(lp, tg, il, tl, red="none", zi=True, blank=0, atol=GRAD_ATOL, from_logits=False):
    import ssak_b200
    x = lp.cuda().requires_grad_(True)
    fn = ssak_b200.ctc_loss_from_logits if from_logits else ssak_b200.ctc_loss
    loss = fn(x, tg, il, tl, blank, red, zi)
    loss.sum().backward()
    y = lp.double().requires_grad_(True)
    ref = F.ctc_loss(F.log_softmax(y, -1) if from_logits else y, tg, il, tl, blank, red, zi)
    ref.sum().backward()
    l, r = loss.detach().cpu().double().reshape(-1), ref.detach().reshape(-1)
    fin = torch.isfinite(r)
    assert torch.equal(torch.isfinite(l), fin)
    if fin.any():
        assert ((l - r).abs() / r.abs().clamp_min(1e-3))[fin].max().item() <= LOSS_RTOL
    g, rg = x.grad.cpu().double(), y.grad
    if torch.isfinite(rg).all():
        assert (g - rg).abs().max().item() <= atol, (g - rg).abs().max().item()
    else:
        assert torch.equal(torch.isnan(g), torch.isnan(rg))
    return loss.detach().cpu(), x.grad.cpu()


@pytest.mark.parametrize("reduction", ["none", "mean", "sum"])
def test_linear_random_and_planted(reduction):
    from ssak_b200.synth import ctc_batch
    for seed, (B, T, V, Lmin, Lmax, planted) in enumerate([(5, 50, 20, 0, 12, False), (7, 120, 50, 5, 40, True),
                                                           (3, 200, 50, 60, 90, True), (4, 64, 1024, 3, 30, False),
                                                           (2, 90, 257, 40, 44, True), (3, 700, 50, 250, 330, True),
                                                           (2, 1100, 30, 480, 511, True), (3, 300, 50, 100, 127, False)]):
        lp, tg, il, tl = ctc_batch(B, T, V, Lmin, Lmax, 800 + seed, Tmin=T // 2, planted=planted)
        _, grad = _check(lp, tg, il, tl, reduction)
        assert (grad[int(il[0]):, 0] == 0).all()


def test_linear_edge_cases():
    g = torch.Generator().manual_seed(5)
    T, B, V = 12, 6, 6
    lp = torch.randn(T, B, V, generator=g).log_softmax(-1)
    tg = torch.tensor([[1, 1, 2, 0, 0], [1, 2, 3, 4, 5], [2, 2, 2, 2, 2], [3, 0, 0, 0, 0], [1, 2, 1, 2, 1], [4, 4, 1, 1, 0]])
    il = torch.tensor([3, 12, 12, 1, 9, 12])           # sample 0: repeated label, too few frames -> inf
    tl = torch.tensor([3, 5, 5, 1, 5, 0])              # sample 5: empty target
    for zi in (True, False):
        for red in ("none", "mean", "sum"):
            _check(lp, tg, il, tl, red, zi)
    # 1, 2, 3 frames; blank != 0; every chunk remainder
    for T2 in (1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 33):
        lp2 = torch.randn(T2, 3, 9, generator=g).log_softmax(-1)
        tg2 = torch.randint(0, 8, (3, 4), generator=g)
        il2 = torch.tensor([T2, max(T2 - 1, 1), max(T2 // 2, 1)])
        tl2 = torch.tensor([min(4, T2), min(2, T2), 1])
        _check(lp2, tg2, il2, tl2, "none", True, blank=8)


def test_linear_hands_back_what_fp64_cannot_hold():
    """-700 log-probabilities (SpeechBrain's padding, speechbrain_infer.py:237-242) inside the lengths, -inf
    emissions and un-normalised positive 'log-probabilities': the utterance is flagged and the log-domain kernels
    recompute it -- same numbers as torch either way."""
    from ssak_b200.synth import ctc_batch
    lp, tg, il, tl = ctc_batch(5, 90, 12, 5, 20, 820, Tmin=70)
    lp[40:48, 1, :] = -700.0
    lp[40:48, 1, 0] = 0.0
    lp[10, 2, 3] = float("-inf")
    _check(lp, tg, il, tl, "none", True, atol=1e-4)
    lp2 = lp.clone()
    lp2[:, 3] = lp2[:, 3] * 0.5 + 1.0                    # not log-probabilities at all: torch accepts them, so do we
    _check(lp2, tg, il, tl, "none", True, atol=1e-4)


def test_linear_from_logits_and_host_abi():
    import ctypes as C
    import ssak_b200
    from ssak_b200.synth import ctc_batch
    g = torch.Generator().manual_seed(7)
    lp, tg, il, tl = ctc_batch(5, 150, 50, 10, 40, 830, Tmin=80)
    logits = lp * 1.7 + 4.0 * torch.randn(150, 5, 1, generator=g) + 3.0
    _check(logits, tg, il, tl, "mean", True, from_logits=True, atol=1e-5)
    L = ssak_b200.lib()
    T, B, V = lp.shape
    lpn, tgn = np.ascontiguousarray(lp.numpy()), np.ascontiguousarray(tg.numpy().astype(np.int32))
    iln, tln = il.numpy().astype(np.int32), tl.numpy().astype(np.int32)
    nll, grad = np.zeros(B, np.float32), np.zeros((T, B, V), np.float32)
    ctx = C.c_void_p()
    assert L.ssak_context_create(0, C.byref(ctx)) == 0
    rc = L.ssak_ctc_loss_host(ctx, lpn.ctypes.data, T, B, V, tgn.ctypes.data, tgn.shape[1], iln.ctypes.data,
                              tln.ctypes.data, 0, 1, None, nll.ctypes.data, grad.ctypes.data)
    L.ssak_context_destroy(ctx)
    assert rc == 0
    y = lp.double().requires_grad_(True)
    ref = F.ctc_loss(y, tg, il, tl, 0, "none", True)
    ref.sum().backward()
    assert np.abs(nll - ref.detach().numpy()).max() <= 1e-5 * np.abs(ref.detach().numpy()).max()
    assert np.abs(grad - y.grad.numpy()).max() <= GRAD_ATOL


def test_linear_workspace_is_small_and_full_size_c2():
    """No stored lattice: the workspace of the C2 batch is ~8x smaller than the log-domain one; full-size parity."""
    import ssak_b200
    from ssak_b200.synth import ctc_batch
    L = ssak_b200.lib()
    lin_ws = L.ssak_ctc_loss_workspace_bytes(1500, 64, 400, 1)
    lp, tg, il, tl = ctc_batch(64, 1500, 50, 200, 400, 1236, Tmin=1200, planted=False)
    idx = [0, 21, 63]
    x = lp.cuda().requires_grad_(True)
    loss = ssak_b200.ctc_loss(x, tg, il, tl, 0, "none", True)
    loss.sum().backward()
    y = lp[:, idx].double().requires_grad_(True)
    ref = F.ctc_loss(y, tg[idx], il[idx], tl[idx], 0, "none", True)
    ref.sum().backward()
    assert ((loss[idx].cpu().double() - ref.detach()).abs() / ref.detach().abs()).max() <= LOSS_RTOL
    err = (x.grad[:, idx].cpu().double() - y.grad).abs().max().item()
    print(f"C2 random emissions, linear-domain kernels: gradient error vs fp64 truth {err:.2e}")
    assert err <= GRAD_ATOL
    import os
    os.environ["SSAK_CTC_LINEAR"] = "0"
    try:
        log_ws = L.ssak_ctc_loss_workspace_bytes(1500, 64, 400, 1)
    finally:
        os.environ["SSAK_CTC_LINEAR"] = "1"
    assert lin_ws < log_ws        # (the log-domain rows stay in the layout for the utterances handed back)
