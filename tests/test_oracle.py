"""The CPU oracle against (1) the committed reference-generated fixtures, (2) the reference's own
code executed live when /root/reference exists (build container), (3) torch's CPU ctc_loss."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O
from oracle import ref_extract as R
from conftest import load_cases


def test_align_oracle_matches_golden(golden_dir):
    cases = load_cases(os.path.join(golden_dir, "align_golden.npz"))
    assert len(cases) >= 40
    n_fail = 0
    for c in cases:
        e, toks = c["emission"], c["tokens"].tolist()
        blank, fag = int(c["blank"]), bool(c["first_as_garbage"])
        tr = O.get_trellis(e, toks, blank, fag)
        assert np.array_equal(tr.view(np.int32), c["trellis"].view(np.int32)), "trellis bits differ"
        if int(c["status"]) != 0:
            n_fail += 1
            with pytest.raises(O.AlignmentFailure):
                O.backtrack(tr, e, toks, blank)
            continue
        path = O.backtrack(tr, e, toks, blank)
        assert [p.token_index for p in path] == c["path_token"].tolist()
        assert [p.time_index for p in path] == c["path_time"].tolist()
        np.testing.assert_allclose([p.score for p in path], c["path_score"], rtol=1e-6, atol=0)
        segs = O.merge_repeats(path)
        assert [s.start for s in segs] == c["seg_start"].tolist()
        assert [s.end for s in segs] == c["seg_end"].tolist()
        np.testing.assert_allclose([s.score for s in segs], c["seg_score"], rtol=1e-6)
        rc, ss, se, sc, ts = O.align(e, toks, blank, fag)
        assert rc == 0 and ts == int(c["t_start"]) and ss.tolist() == c["seg_start"].tolist()
    assert n_fail == 2


def test_ctc_oracle_matches_golden(golden_dir):
    cases = load_cases(os.path.join(golden_dir, "ctc_golden.npz"))
    assert len(cases) >= 40
    for c in cases:
        for key, dt, ltol, gtol in (("f32", np.float32, 1e-6, 2e-5), ("f64", np.float64, 1e-12, 1e-12)):
            loss, nll, grad = O.ctc_loss(c["log_probs"], c["targets"], c["input_lengths"], c["target_lengths"],
                                         int(c["blank"]), str(c["reduction"]), bool(c["zero_infinity"]), dtype=dt)
            ref_l, ref_g = c[f"loss_{key}"], c[f"grad_{key}"]
            assert np.array_equal(np.isfinite(np.asarray(loss)), np.isfinite(ref_l))
            fin = np.isfinite(ref_l)
            np.testing.assert_allclose(np.asarray(loss)[fin], ref_l[fin], rtol=ltol)
            if np.isfinite(ref_g).all():
                np.testing.assert_allclose(grad, ref_g, atol=gtol, rtol=0)


def test_greedy_oracle_matches_golden(golden_dir):
    for c in load_cases(os.path.join(golden_dir, "greedy_golden.npz")):
        p, rel, blank = c["probs"], c["rel_lens"], int(c["blank"])
        B, T, V = p.shape
        for b in range(B):
            n = int(torch.round(torch.tensor(rel[b]) * T).item())
            ids, fid = O.greedy(p[b], n, blank)
            assert ids == c["out"][b, : c["out_lens"][b]].tolist()
            assert fid.tolist() == c["argmax"][b, :n].tolist()


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_align_oracle_matches_live_reference():
    rng = np.random.default_rng(5)
    for trial in range(120):
        g = torch.Generator().manual_seed(3000 + trial)
        T, V = int(rng.integers(1, 70)), int(rng.integers(2, 12))
        L = int(rng.integers(0, min(T + 3, 30)))
        blank = int(rng.integers(0, V)) if trial % 3 == 0 else 0
        toks = rng.integers(0, V, size=L).tolist()
        fag = trial % 5 == 0 and L > 0
        if trial % 3 == 1:
            e = ((torch.round(2 * torch.randn(T, V, generator=g) * 2) / 2) - 8).numpy()
        else:
            e = torch.randn(T, V, generator=g).log_softmax(-1).numpy()
        ref = R.align(e, toks, blank, fag, want_trellis=True)
        tr = O.get_trellis(e, toks, blank, fag)
        assert np.array_equal(tr.view(np.int32), ref["trellis"].view(np.int32))
        rc, ss, se, sc, ts = O.align(e, toks, blank, fag)
        assert rc == ref["status"] and ts == ref["t_start"]
        if rc == 0:
            assert ss.tolist() == [s[1] for s in ref["segments"]]
            assert se.tolist() == [s[2] for s in ref["segments"]]
            np.testing.assert_allclose(sc, [s[3] for s in ref["segments"]], rtol=1e-6)


def test_ctc_oracle_matches_torch_cpu():
    rng = np.random.default_rng(9)
    for trial in range(12):
        T, B, V = int(rng.integers(2, 60)), int(rng.integers(1, 5)), int(rng.integers(3, 20))
        il = rng.integers(1, T + 1, size=B)
        tl = np.array([int(rng.integers(0, min(int(il[b]), 15) + 1)) for b in range(B)])
        tg = rng.integers(1, V, size=(B, max(int(tl.max()), 1)))
        lp = torch.randn(T, B, V).log_softmax(-1)
        x = lp.double().requires_grad_(True)
        loss = F.ctc_loss(x, torch.tensor(tg), torch.tensor(il), torch.tensor(tl), reduction="sum", zero_infinity=True)
        loss.backward()
        lo, _, g = O.ctc_loss(lp.numpy(), tg, il, tl, 0, "sum", True, dtype=np.float64)
        np.testing.assert_allclose(float(lo), float(loss), rtol=1e-12)
        np.testing.assert_allclose(g, x.grad.numpy(), atol=1e-12)


def test_ctc_oracle_argument_errors():
    lp = np.zeros((4, 1, 3), np.float32)
    with pytest.raises(RuntimeError):
        O.ctc_loss(lp, [[1]], [5], [1])           # input_length > T
    with pytest.raises(RuntimeError):
        O.ctc_loss(lp, [[1]], [4], [1], blank=3)   # blank outside the vocabulary


def test_ctc_oracle_tight_alignments():
    """The shapes the GPU fuzz found the kernels wrong on (tests/test_gpu_tight.py): targets nearly as long as the
    input under peaky emissions that disagree with them, likelihoods of hundreds of nats.  The checker itself must
    be right there: C restatement in fp64 against torch's CPU kernel in fp64, and in fp32 as close to the fp64
    truth as torch's own fp32 kernel is."""
    from ssak_b200.synth import ctc_batch
    for case in ("33 121 122 8 1 0 971909417", "50 205 186 3 1 0 29280320", "33 256 147 3 1 0 224398287"):
        V, Lmax, T, B, planted, _, seed = (int(v) for v in case.split())
        lp, tg, il, tl = ctc_batch(B, T, V, 0, Lmax, seed, Tmin=1, planted=bool(planted))
        tl = torch.minimum(tl, torch.tensor(Lmax))
        il = torch.clamp(il, 1, T)
        x = lp.double().requires_grad_(True)
        ref = F.ctc_loss(x, tg, il, tl, 0, "none", True)
        ref.sum().backward()
        assert float(ref.detach().max()) > 100.0             # (the case is what it claims to be)
        _, nll, g = O.ctc_loss(lp.numpy(), tg.numpy(), il.numpy(), tl.numpy(), 0, "sum", True, dtype=np.float64)
        fin = np.isfinite(nll)
        np.testing.assert_allclose(np.where(fin, nll, 0.0), ref.detach().numpy(), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(g, x.grad.numpy(), atol=1e-11)
        x32 = lp.clone().requires_grad_(True)
        F.ctc_loss(x32, tg, il, tl, 0, "sum", True).backward()
        _, _, g32 = O.ctc_loss(lp.numpy(), tg.numpy(), il.numpy(), tl.numpy(), 0, "sum", True, dtype=np.float32)
        err_o = np.abs(g32.astype(np.float64) - x.grad.numpy()).max()
        err_t = (x32.grad.double() - x.grad).abs().max().item()
        assert err_o <= max(1e-4, 2.0 * err_t + 1e-5), (err_o, err_t)
