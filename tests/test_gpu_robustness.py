"""Behaviour at the edges of the contract (round-1 review): NaN likelihoods stay visible under zero_infinity,
labels outside the vocabulary are errors (never clamped into a wrong number), shapes the kernels do not cover are
delegated by install(), concurrent aligner calls on two streams of one GPU, and the C-ABI status codes."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_nan_emissions_stay_nan_with_zero_infinity():
    """torch: where(loss == inf, 0, loss) keeps a NaN; so does the gradient of that utterance.  (The reference's HF
    recipe trains with zero_infinity=True, wav2vec_train.py:325: a diverged model must not be masked.)"""
    import ssak_b200
    from ssak_b200.synth import ctc_batch
    lp, tg, il, tl = ctc_batch(4, 60, 20, 3, 12, 900, Tmin=40)
    lp[7, 1, :] = float("nan")
    tg[2, :3] = torch.tensor([5, 5, 5])
    il[2], tl[2] = 4, 3                      # infeasible -> +inf -> zeroed
    x = lp.cuda().requires_grad_(True)
    loss = ssak_b200.ctc_loss(x, tg, il, tl, 0, "none", True)
    loss.sum().backward()
    y = lp.double().requires_grad_(True)
    ref = F.ctc_loss(y, tg, il, tl, 0, "none", True)
    ref.sum().backward()
    l = loss.detach().cpu()
    assert torch.isnan(l[1]) and torch.isnan(ref[1]), (l, ref)
    assert l[2] == 0 and ref[2] == 0
    assert torch.isnan(x.grad[: int(il[1]), 1].cpu()).all()
    assert (x.grad[:, 2] == 0).all()
    for b in (0, 3):
        assert abs(l[b].item() - ref[b].item()) <= 1e-5 * abs(ref[b].item())
        assert (x.grad[:, b].cpu().double() - y.grad[:, b]).abs().max() <= 1e-4
    red = ssak_b200.ctc_loss(lp.cuda(), tg, il, tl, 0, "mean", True)
    assert torch.isnan(red)


def test_out_of_range_targets_are_errors():
    import ssak_b200
    from ssak_b200.synth import ctc_batch
    lp, tg, il, tl = ctc_batch(3, 50, 10, 3, 8, 901, Tmin=30)
    bad = tg.clone()
    bad[1, 0] = 10                           # == V
    with pytest.raises(RuntimeError, match="label range"):
        ssak_b200.ctc_loss(lp.cuda(), bad, il, tl)
    # device targets cannot be checked without a sync: the utterance's likelihood and gradient are NaN, others intact
    x = lp.cuda().requires_grad_(True)
    loss = ssak_b200.ctc_loss(x, bad.cuda(), il.cuda(), tl.cuda(), 0, "none", True)
    loss.sum().backward()
    ok = ssak_b200.ctc_loss(lp.cuda(), tg, il, tl, 0, "none", True)
    assert torch.isnan(loss[1]) and torch.isnan(x.grad[: int(il[1]), 1]).all()
    assert torch.equal(loss[[0, 2]], ok[[0, 2]])
    # host-buffer C entry point: SSAK_ERR_INVALID_ARGUMENT
    L = ssak_b200.lib()
    ctx = C.c_void_p()
    assert L.ssak_context_create(0, C.byref(ctx)) == 0
    lpn, tgn = np.ascontiguousarray(lp.numpy()), np.ascontiguousarray(bad.numpy().astype(np.int32))
    iln, tln = il.numpy().astype(np.int32), tl.numpy().astype(np.int32)
    nll = np.zeros(3, np.float32)
    rc = L.ssak_ctc_loss_host(ctx, lpn.ctypes.data, 50, 3, 10, tgn.ctypes.data, tgn.shape[1], iln.ctypes.data,
                              tln.ctypes.data, 0, 1, None, nll.ctypes.data, None)
    assert rc == -1
    L.ssak_context_destroy(ctx)
    # aligner: a token outside the vocabulary -> status 3 for that utterance only
    from ssak_b200.synth import align_batch
    em, toks, el, tl2 = align_batch(3, 80, 12, 5, 20, 902, Tmin=60)
    toks[2, 1] = 99
    res = ssak_b200.forced_align(em.cuda(), toks, el, tl2)
    assert res.status.cpu().tolist() == [0, 0, 3]
    assert (res.starts[2] == -1).all()


def test_install_delegates_unsupported_shapes():
    """V = 4096 (a large BPE vocabulary) does not fit the emission ring: after install() F.ctc_loss must still
    work (torch's own CUDA kernel), not raise."""
    import ssak_b200
    assert ssak_b200.loss.supported(100, 4, 1024, 200) and not ssak_b200.loss.supported(100, 4, 4096, 20)
    assert not ssak_b200.loss.supported(100, 4, 50, 5000)
    g = torch.Generator().manual_seed(3)
    lp = torch.randn(40, 2, 4096, generator=g).log_softmax(-1)
    tg = torch.randint(1, 4096, (2, 7), generator=g)
    il, tl = torch.tensor([40, 33]), torch.tensor([7, 5])
    ref = F.ctc_loss(lp.double(), tg, il, tl, 0, "mean", True)
    ssak_b200.install()
    try:
        out = F.ctc_loss(lp.cuda(), tg.cuda(), il, tl, 0, "mean", True)
        small = F.ctc_loss(lp[:, :, :50].log_softmax(-1).cuda(), tg.clamp_max(49).cuda(), il, tl, 0, "mean", True)
    finally:
        ssak_b200.uninstall()
    assert abs(out.item() - ref.item()) <= 1e-4 * abs(ref.item())
    assert torch.isfinite(small)
    with pytest.raises(ssak_b200.SsakB200Error, match="V=4096"):
        ssak_b200.ctc_loss(lp.cuda(), tg, il, tl)


def test_two_concurrent_aligner_calls_on_two_streams():
    """Two multi-CTA alignments (S > 1: CTAs of an utterance poll each other's seams) issued back to back on two
    streams of one GPU: both complete with status 0 and oracle-exact spans (cooperative / cluster launch)."""
    import ssak_b200
    from ssak_b200.synth import align_batch
    batches = [align_batch(12, 2400, 50, 1500, 1800, 910 + i, Tmin=2300) for i in range(2)]
    dev = torch.device("cuda")
    ems = [b[0].to(dev) for b in batches]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    results = [None, None]
    for rep in range(3):
        for i in (0, 1):
            with torch.cuda.stream(streams[i]):
                results[i] = ssak_b200.forced_align(ems[i], batches[i][1], batches[i][2], batches[i][3])
    torch.cuda.synchronize()
    for i in (0, 1):
        em, toks, el, tl = batches[i]
        res = results[i]
        assert (res.status == 0).all(), res.status
        st, en = res.starts.cpu().numpy(), res.ends.cpu().numpy()
        for b in (0, 5, 11):
            Tb, Lb = int(el[b]), int(tl[b])
            rc, ss, se, _, t0 = O.align(em[b, :Tb].numpy(), toks[b, :Lb].tolist(), 0, False)
            assert rc == 0 and int(res.t_start[b]) == t0
            assert st[b, :Lb].tolist() == ss.tolist() and en[b, :Lb].tolist() == se.tolist()


def test_two_host_threads_different_shapes():
    """The header promises re-entrancy from several host threads: two threads, different loss shapes (different
    shared-memory sizes of the same kernel instantiation), many iterations."""
    import threading
    import ssak_b200
    from ssak_b200.synth import ctc_batch
    shapes = [(6, 200, 50, 20, 60), (5, 260, 50, 90, 120)]
    data = [ctc_batch(B, T, V, lo, hi, 920 + i, Tmin=T - 20) for i, (B, T, V, lo, hi) in enumerate(shapes)]
    refs = [F.ctc_loss(d[0].double(), d[1], d[2], d[3], 0, "none", True) for d in data]
    errs = []

    def work(i):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            lp, tg, il, tl = data[i]
            x = lp.cuda()
            with torch.cuda.stream(st):
                for _ in range(40):
                    y = x.detach().requires_grad_(True)
                    loss = ssak_b200.ctc_loss(y, tg, il, tl, 0, "none", True)
                    loss.sum().backward()
                st.synchronize()
            rel = ((loss.detach().cpu().double() - refs[i]).abs() / refs[i].abs()).max().item()
            if rel > 1e-5:
                errs.append((i, rel))
        except Exception as e:   # noqa: BLE001
            errs.append((i, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
