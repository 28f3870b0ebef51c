"""BASELINE config C1: the reference's tests/data Kaldi folder through a random-init wav2vec2-base CTC head
(tests/golden/make_c1.py ran the model, the reference aligner and torch's CPU ctc_loss in the build
container; the audio is not needed here).  CPU: the oracle reproduces the fixture.  GPU: the kernels do."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

BLANK = 0


def _load(golden_dir):
    z = np.load(os.path.join(golden_dir, "c1_golden.npz"), allow_pickle=False)
    return z, int(z["n"])


def test_c1_oracle_matches_reference(golden_dir):
    z, n = _load(golden_dir)
    assert n >= 4
    for i in range(n):
        em, toks = z[f"u{i}_emission"], z[f"u{i}_tokens"].tolist()
        for fag in (0, 1):
            rc, ss, se, sc, ts = O.align(em, toks, BLANK, bool(fag))
            k = f"u{i}_g{fag}_"
            assert rc == int(z[k + "status"]) and ts == int(z[k + "t_start"])
            if rc == 0:
                assert ss.tolist() == z[k + "seg_start"].tolist() and se.tolist() == z[k + "seg_end"].tolist()
                np.testing.assert_allclose(sc, z[k + "seg_score"], rtol=1e-6)
    loss, _, grad = O.ctc_loss(z["log_probs"], z["targets"], z["input_lengths"], z["target_lengths"], BLANK, "mean",
                               True, dtype=np.float64)
    np.testing.assert_allclose(float(loss), float(z["loss_f64"]), rtol=1e-12)
    np.testing.assert_allclose(grad, z["grad_f64"], atol=1e-12)


@pytest.mark.gpu
def test_c1_kernels_match_reference(golden_dir):
    import ssak_b200
    z, n = _load(golden_dir)
    ems = [torch.from_numpy(z[f"u{i}_emission"]) for i in range(n)]
    toks = [z[f"u{i}_tokens"] for i in range(n)]
    Tmax, Lmax, V = max(e.shape[0] for e in ems), max(len(t) for t in toks), ems[0].shape[1]
    em = torch.zeros(n, Tmax, V)
    tk = torch.zeros(n, Lmax, dtype=torch.int32)
    el, tl = torch.zeros(n, dtype=torch.int32), torch.zeros(n, dtype=torch.int32)
    for i in range(n):
        em[i, : ems[i].shape[0]] = ems[i]
        tk[i, : len(toks[i])] = torch.from_numpy(toks[i])
        el[i], tl[i] = ems[i].shape[0], len(toks[i])
    for fag in (0, 1):
        kw = {}
        if fag:  # column 0 with the reference's CPU ops (see test_gpu_align.test_align_golden_fixtures)
            kw["col0"] = torch.stack([(1 - em[i, :, int(tk[i, 0])].exp()).log() for i in range(n)])
        res = ssak_b200.forced_align(em.cuda(), tk, el, tl, blank_id=BLANK, first_as_garbage=bool(fag), **kw)
        for i in range(n):
            k = f"u{i}_g{fag}_"
            L = int(tl[i])
            assert int(res.status[i]) == (0 if int(z[k + "status"]) == 0 else 1)
            if int(z[k + "status"]) == 0:
                assert int(res.t_start[i]) == int(z[k + "t_start"])
                assert res.starts[i, :L].cpu().tolist() == z[k + "seg_start"].tolist()
                assert res.ends[i, :L].cpu().tolist() == z[k + "seg_end"].tolist()
                np.testing.assert_allclose(res.scores[i, :L].cpu().numpy(), z[k + "seg_score"], rtol=1e-6)
    # the HF loss call on the same batch
    x = torch.from_numpy(z["log_probs"]).cuda().requires_grad_(True)
    loss = ssak_b200.ctc_loss(x, torch.from_numpy(z["targets"]), torch.from_numpy(z["input_lengths"]),
                              torch.from_numpy(z["target_lengths"]), blank=BLANK, reduction="mean", zero_infinity=True)
    loss.backward()
    assert abs(loss.item() - float(z["loss_f64"])) <= 1e-5 * abs(float(z["loss_f64"]))
    assert np.abs(x.grad.cpu().numpy() - z["grad_f64"]).max() <= 1e-4
    assert np.abs(x.grad.cpu().numpy() - z["grad_f32"]).max() <= 1e-4
    # greedy decode of the same emissions = argmax + group-by + drop blank
    out = ssak_b200.ctc_greedy_decode(em.cuda(), el.float() / Tmax, blank_id=BLANK)
    for i in range(n):
        assert out[i] == O.greedy(em[i].numpy(), int(el[i]), BLANK)[0]
