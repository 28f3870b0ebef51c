"""The throughput loss kernels (ssak_b200/csrc/ctc_lin32.cu: one warp per (utterance, direction), linear domain,
block floating point, no stored lattice) -- the default path for V <= 128 and targets up to 415 labels.  Same results
as torch's CPU ctc_loss in fp64, closer to it than any fp32 log-domain recursion; utterances whose self-check fails
(garbage transcripts, -700 padding, NaN / inf emissions, likelihoods ~1) are recomputed by the log-domain kernels in
the same call and still match."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_lin32(monkeypatch):
    """Small batches default to the latency-tuned log-domain kernels (a chain is ONE warp here): force the
    throughput kernels wherever they are valid.  test_gpu_timed_configs.py covers the default dispatch at B = 1024."""
    monkeypatch.setenv("SSAK_CTC_LIN32", "1")


LOSS_RTOL = 1e-5
GRAD_ATOL = 1e-5          # block floating point rounds relatively: ~2e-6 at T = 1500 (tools/proto_bfp.py)


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
    assert torch.equal(torch.isfinite(l), fin), (l, r)
    if fin.any():
        rel = ((l - r).abs() / r.abs().clamp_min(1.0))[fin].max().item()   # 1e-5 relative (absolute below 1)
        assert rel <= LOSS_RTOL, rel
    g, rg = x.grad.cpu().double(), y.grad
    # torch forms exp(lcab - lp + nll) and gets (-inf) - (-inf) = NaN exactly where an emission the target uses is -inf;
    # the true gradient there is 0, which is what the kernels return (the one documented deviation)
    art = torch.isneginf(lp) & torch.isnan(rg) if not from_logits else torch.zeros_like(rg, dtype=torch.bool)
    assert torch.equal(torch.isnan(g) & ~art, torch.isnan(rg) & ~art)
    assert (g[art] == 0).all() or torch.isnan(g[art]).all()
    ok = ~torch.isnan(rg)
    err = (g - rg)[ok].abs().max().item() if ok.any() else 0.0
    assert err <= atol, err
    return loss.detach().cpu(), x.grad.cpu()


def _path_flags(lp, tg, il, tl, backward=True, blank=0):
    """Run the C ABI directly on our own workspace and return the per-utterance path flags."""
    import ssak_b200
    L = ssak_b200.lib()
    T, B, V = lp.shape
    dev = torch.device("cuda", 0)
    x = lp.to(dev).contiguous()
    tg32 = tg.to(dev, torch.int32).contiguous()
    off = torch.arange(B, device=dev, dtype=torch.int64) * tg32.shape[1]
    il32, tl32 = il.to(dev, torch.int32), tl.to(dev, torch.int32)
    lmax = int(tl.max())
    wsb = L.ssak_ctc_loss_workspace_bytes_v(T, B, V, lmax, 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    nll = torch.empty(B, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    rc = L.ssak_ctc_loss_forward(x.data_ptr(), T, B, V, x.stride(0), x.stride(1), tg32.data_ptr(), off.data_ptr(),
                                 il32.data_ptr(), tl32.data_ptr(), lmax, blank, 1, nll.data_ptr(), ws.data_ptr(), wsb, s)
    assert rc == 0
    grad = torch.empty_like(x)
    if backward:
        go = torch.ones(B, device=dev)
        rc = L.ssak_ctc_loss_backward(go.data_ptr(), x.data_ptr(), T, B, V, x.stride(0), x.stride(1), tg32.data_ptr(),
                                      off.data_ptr(), il32.data_ptr(), tl32.data_ptr(), lmax, blank, 1, nll.data_ptr(),
                                      grad.data_ptr(), grad.stride(0), grad.stride(1), ws.data_ptr(), wsb, s)
        assert rc == 0
    fl = torch.empty(B, dtype=torch.int32, device=dev)
    assert L.ssak_ctc_loss_path_flags(ws.data_ptr(), T, B, V, lmax, 1, fl.data_ptr(), s) == 0
    torch.cuda.synchronize()
    return fl.cpu(), nll.cpu(), grad.cpu()


@pytest.mark.parametrize("reduction", ["none", "mean", "sum"])
def test_lin32_random_and_planted(reduction):
    from ssak_b200.synth import ctc_batch
    for seed, (B, T, V, Lmin, Lmax, planted) in enumerate([(5, 50, 20, 0, 12, False), (7, 120, 50, 5, 40, True),
                                                           (3, 200, 50, 60, 90, True), (4, 64, 128, 3, 30, False),
                                                           (2, 90, 97, 40, 44, True), (3, 700, 50, 250, 330, True),
                                                           (2, 1100, 30, 400, 415, True), (3, 300, 50, 100, 127, False),
                                                           (3, 300, 33, 128, 223, False), (40, 150, 50, 20, 60, True),
                                                           # large vocabularies: the gather kernels (ctc_lin32_lv.cuh)
                                                           (4, 64, 1024, 3, 30, False), (3, 200, 260, 60, 90, True),
                                                           (2, 500, 2000, 150, 223, True), (5, 120, 200, 5, 40, True),
                                                           (3, 90, 132, 100, 127, False)]):
        lp, tg, il, tl = ctc_batch(B, T, V, Lmin, Lmax, 800 + seed, Tmin=T // 2, planted=planted)
        _, grad = _check(lp, tg, il, tl, reduction)
        assert (grad[int(il[0]):, 0] == 0).all()


def test_lin32_is_the_path_that_ran():
    """Ordinary batches stay on the throughput kernels (flags 0): a silent hand-back would only show as lost speed."""
    from ssak_b200.synth import ctc_batch
    for planted in (True, False):
        lp, tg, il, tl = ctc_batch(6, 400, 50, 60, 120, 77, Tmin=300, planted=planted)
        fl, _, _ = _path_flags(lp, tg, il, tl)
        assert (fl == 0).all(), fl
    # early training: nearly uniform emissions (the diffusive regime the per-lane exponents exist for)
    g = torch.Generator().manual_seed(3)
    lp = (0.01 * torch.randn(1500, 3, 50, generator=g)).log_softmax(-1)
    tg = torch.randint(1, 50, (3, 400), generator=g)
    il, tl = torch.tensor([1500, 1400, 1300]), torch.tensor([400, 300, 200])
    fl, _, _ = _path_flags(lp, tg, il, tl)
    assert (fl == 0).all(), fl
    _check(lp, tg, il, tl)


def test_lin32_edge_cases():
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
    for T2 in (1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 33):
        lp2 = torch.randn(T2, 3, 9, generator=g).log_softmax(-1)
        tg2 = torch.randint(0, 8, (3, 4), generator=g)
        il2 = torch.tensor([T2, max(T2 - 1, 1), max(T2 // 2, 1)])
        tl2 = torch.tensor([min(4, T2), min(2, T2), 1])
        _check(lp2, tg2, il2, tl2, "none", True, blank=8)
    # strided input (HF's transposed view) and a strided gradient
    import ssak_b200
    lpb = torch.randn(5, 40, 20, generator=g).log_softmax(-1)      # [B,T,V]
    tg3 = torch.randint(1, 20, (5, 9), generator=g)
    il3, tl3 = torch.tensor([40, 33, 21, 40, 19]), torch.tensor([9, 4, 7, 1, 9])
    x = lpb.cuda().requires_grad_(True)
    ssak_b200.ctc_loss(x.transpose(0, 1), tg3, il3, tl3, 0, "sum", True).backward()
    y = lpb.double().requires_grad_(True)
    F.ctc_loss(y.transpose(0, 1), tg3, il3, tl3, 0, "sum", zero_infinity=True).backward()
    assert (x.grad.cpu().double() - y.grad).abs().max().item() <= GRAD_ATOL


def test_lin32_hands_back_what_fp32_cannot_hold():
    """-700 log-probabilities (SpeechBrain's padding, speechbrain_infer.py:237-242) inside the lengths, -inf
    emissions, NaN emissions and un-normalised positive 'log-probabilities': the utterance is flagged and the
    log-domain kernels recompute it -- same numbers as torch either way."""
    from ssak_b200.synth import ctc_batch
    lp, tg, il, tl = ctc_batch(5, 90, 12, 5, 20, 820, Tmin=70)
    lp[40:48, 1, :] = -700.0
    lp[40:48, 1, 0] = 0.0
    lp[10, 2, 3] = float("-inf")
    _check(lp, tg, il, tl, "none", True, atol=1e-4)
    lp2 = lp.clone()
    lp2[:, 3] = lp2[:, 3] * 0.5 + 1.0                    # not log-probabilities at all: torch accepts them, so do we
    _check(lp2, tg, il, tl, "none", True, atol=1e-4)
    fl, _, _ = _path_flags(lp2, tg, il, tl)
    assert fl[3] != 0 and fl[0] == 0, fl
    lp3 = lp.clone()
    lp3[20, 4, :] = float("nan")
    _check(lp3, tg, il, tl, "none", True, atol=1e-4)
    _check(lp3, tg, il, tl, "mean", False, atol=1e-4)


def _garbage_batch(boost, ND, seed=11, T=1500, V=50, L=300, B=3):
    """Peaky planted emissions; utterance 1's transcript holds ND labels the audio does not contain."""
    g = torch.Generator().manual_seed(seed)
    tg = torch.randint(1, V, (B, L), generator=g)
    lp = torch.empty(T, B, V)
    for b in range(B):
        lg = torch.randn(T, V, generator=g)
        lg[:, 0] += boost
        keep = [i for i in range(L) if not (b == 1 and 100 <= i < 100 + ND)]
        onset = torch.sort(torch.randperm(T, generator=g)[: len(keep)]).values
        lg[onset, 0] -= boost
        lg[onset, tg[b, keep]] += boost
        lp[:, b] = lg.log_softmax(-1)
    return lp, tg, torch.full((B,), T), torch.full((B,), L)


@pytest.mark.parametrize("boost,ND", [(6.0, 25), (12.0, 60)])
def test_lin32_garbage_transcript(boost, ND):
    """A transcript with labels the audio does not contain: the forward and the backward partial likelihoods of one
    lane disagree by up to 2^(17 ND).  Either the block floating point holds it (flags 0) or the per-frame mass check
    of backward() catches it and the log-domain kernels take over in the same call -- same gradient as torch either
    way; the harsher case must be handed back."""
    lp, tg, il, tl = _garbage_batch(boost, ND)
    fl, nll, grad = _path_flags(lp, tg, il, tl)
    y = lp.double().requires_grad_(True)
    ref = F.ctc_loss(y, tg, il, tl, 0, "none", True)
    ref.sum().backward()
    err = (grad.double() - y.grad).abs().amax(dim=(0, 2))
    print(f"boost {boost} ND {ND}: flags {fl.tolist()} gradient error per utterance {err.tolist()}")
    assert fl[0] == 0 and fl[2] == 0 and not (fl[1] & 4), fl
    assert ((nll.double() - ref.detach()).abs() / ref.detach().abs()).max().item() <= LOSS_RTOL
    assert err[0] <= GRAD_ATOL and err[2] <= GRAD_ATOL
    # (a handed-back garbage transcript gets the fp32 log-domain recursion's accuracy on a loss of ~1e4)
    assert err[1] <= (GRAD_ATOL if fl[1] == 0 else 1e-3)
    if boost > 10:
        assert fl[1] & 3, fl


def test_lin32_tiny_loss():
    """A likelihood close to 1 (loss ~0.02): fp32 states carry log P to ~1e-6 absolute (any fp32 recursion does)."""
    T, V, L = 200, 30, 20
    g = torch.Generator().manual_seed(2)
    tg = torch.randint(1, V, (2, L), generator=g)
    lg = torch.zeros(T, 2, V)
    lg[:, :, 0] = 14.0
    onset = torch.arange(L) * 9 + 3
    for b in range(2):
        lg[onset, b, 0] = 0.0
        lg[onset, b, tg[b]] = 14.0
    lp = lg.log_softmax(-1)
    il, tl = torch.tensor([T, T]), torch.tensor([L, L])
    fl, _, _ = _path_flags(lp, tg, il, tl)
    assert (fl == 0).all(), fl
    _check(lp, tg, il, tl, "none", True)


def test_lin32_more_hand_backs_than_row_blocks_is_loud():
    """Row blocks for handed-back utterances exist for max(32, B/8) of them: beyond that the likelihood and the gradient
    are NaN (never a silently wrong number); the ones that got a block still match torch."""
    from ssak_b200.synth import ctc_batch
    lp, tg, il, tl = ctc_batch(40, 60, 12, 3, 10, 905, Tmin=40)
    lp = lp * 0.5 + 1.0                                   # 'log-probabilities' above 0: every utterance is handed back
    fl, nll, grad = _path_flags(lp, tg, il, tl)
    assert int((fl & 4).ne(0).sum()) == 8 and int((fl == 1).sum()) == 32, fl
    y = lp.double().requires_grad_(True)
    ref = F.ctc_loss(y, tg, il, tl, 0, "none", True)
    ref.sum().backward()
    for b in range(40):
        if fl[b] & 4:
            assert torch.isnan(nll[b]) and torch.isnan(grad[: int(il[b]), b]).all() and (grad[int(il[b]):, b] == 0).all()
        else:
            assert abs(float(nll[b]) - float(ref[b])) <= 1e-4 * abs(float(ref[b]))
            assert (grad[:, b].double() - y.grad[:, b]).abs().max() <= 1e-3


def test_log_domain_kernels_still_default_elsewhere_and_on_request(monkeypatch):
    """V > 128 or targets beyond 415 labels stay on the log-domain kernels; SSAK_CTC_LIN32=0 forces them."""
    from ssak_b200.synth import ctc_batch
    lp, tg, il, tl = ctc_batch(3, 120, 200, 5, 40, 901, Tmin=80)
    _check(lp, tg, il, tl, "none", True, atol=1e-4)
    lp, tg, il, tl = ctc_batch(2, 1300, 20, 500, 600, 902, Tmin=1250)
    _check(lp, tg, il, tl, "none", True, atol=1e-4)
    monkeypatch.setenv("SSAK_CTC_LIN32", "0")
    lp, tg, il, tl = ctc_batch(4, 300, 50, 20, 100, 903, Tmin=200)
    _check(lp, tg, il, tl, "mean", True, atol=1e-4)
    _check(lp * 1.3 + 0.7, tg, il, tl, "mean", True, atol=1e-4, from_logits=True)


def test_lin32_from_logits_and_host_abi():
    import ssak_b200
    from ssak_b200.synth import ctc_batch
    g = torch.Generator().manual_seed(7)
    lp, tg, il, tl = ctc_batch(5, 150, 50, 10, 40, 830, Tmin=80)
    logits = lp * 1.7 + 4.0 * torch.randn(150, 5, 1, generator=g) + 3.0
    _check(logits, tg, il, tl, "mean", True, from_logits=True)
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


def test_lin32_large_vocabulary_logits_strides_and_hand_back():
    """The gather kernels (V > 128): logits entry point, HF's transposed view (strided rows: scalar gradient pass),
    an utterance that has to be handed back."""
    import ssak_b200
    from ssak_b200.synth import ctc_batch
    g = torch.Generator().manual_seed(9)
    lp, tg, il, tl = ctc_batch(4, 140, 512, 10, 60, 840, Tmin=90)
    logits = lp * 1.3 + 2.0 * torch.randn(140, 4, 1, generator=g) - 1.0
    _check(logits, tg, il, tl, "mean", True, from_logits=True)
    lpb = torch.randn(3, 70, 300, generator=g).log_softmax(-1)      # [B,T,V]
    tg3 = torch.randint(1, 300, (3, 20), generator=g)
    il3, tl3 = torch.tensor([70, 51, 33]), torch.tensor([20, 7, 12])
    x = lpb.cuda().requires_grad_(True)
    ssak_b200.ctc_loss(x.transpose(0, 1), tg3, il3, tl3, 0, "sum", True).backward()
    y = lpb.double().requires_grad_(True)
    F.ctc_loss(y.transpose(0, 1), tg3, il3, tl3, 0, "sum", zero_infinity=True).backward()
    assert (x.grad.cpu().double() - y.grad).abs().max().item() <= GRAD_ATOL
    lp2 = lp.clone()
    lp2[:, 2] = lp2[:, 2] * 0.5 + 1.0                                # not log-probabilities: handed back
    _check(lp2, tg, il, tl, "none", True, atol=1e-4)
    fl, _, _ = _path_flags(lp2, tg, il, tl)
    assert fl[2] != 0 and fl[0] == 0 and fl[1] == 0 and fl[3] == 0, fl
    fl, _, _ = _path_flags(lp, tg, il, tl)
    assert (fl == 0).all(), fl


def test_lin32_workspace_is_small_and_full_size_c2():
    """No stored lattice: the workspace of the C2 batch is several times smaller than the log-domain one; full-size
    parity on random emissions (the case where fp32 log-domain recursions are 1e-4 .. 3e-3 off)."""
    import ssak_b200
    from ssak_b200.synth import ctc_batch
    L = ssak_b200.lib()
    lin_ws = L.ssak_ctc_loss_workspace_bytes_v(1500, 64, 50, 400, 1)
    log_ws = L.ssak_ctc_loss_workspace_bytes_v(1500, 64, 200, 400, 1)
    assert lin_ws < log_ws and L.ssak_ctc_loss_workspace_bytes(1500, 64, 400, 1) >= log_ws, (lin_ws, log_ws)
    assert 2.5 * L.ssak_ctc_loss_workspace_bytes_v(1500, 1024, 50, 400, 1) < L.ssak_ctc_loss_workspace_bytes_v(1500, 1024, 200, 400, 1)
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
    print(f"C2 random emissions, throughput kernels: gradient error vs fp64 truth {err:.2e}")
    assert err <= GRAD_ATOL
