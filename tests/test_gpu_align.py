"""Parity of the CUDA aligner (through the C ABI) with the oracle / the reference-generated
fixtures.  Bar: spans, t_start, status and the decision path bit-exact; per-token scores (they
contain an exp) within 1e-6 relative."""
import os

import numpy as np
import pytest
import torch

from conftest import load_cases
from oracle import oracle as O

pytestmark = pytest.mark.gpu
SCORE_RTOL = 1e-6


def _run(em, toks, el=None, tl=None, blank=0, fag=False, **kw):
    import ssak_b200
    res = ssak_b200.forced_align(em.cuda(), toks, el, tl, blank_id=blank, first_as_garbage=fag, **kw)
    torch.cuda.synchronize()
    return res


def _check_batch(em, toks, el, tl, blank=0, fag=False, res=None, tag=""):
    res = _run(em, toks, el, tl, blank, fag) if res is None else res
    st, en, sc = res.starts.cpu().numpy(), res.ends.cpu().numpy(), res.scores.cpu().numpy()
    status, ts = res.status.cpu().numpy(), res.t_start.cpu().numpy()
    n_ok = 0
    for b in range(em.shape[0]):
        Tb, Lb = int(el[b]), int(tl[b])
        rc, ss, se, ssc, t0 = O.align(em[b, :Tb].numpy(), toks[b, :Lb].tolist(), blank, fag)
        assert (status[b] == 0) == (rc == 0), f"{tag} b={b} status {status[b]} vs oracle {rc}"
        if rc != 0:
            continue
        n_ok += 1
        assert ts[b] == t0, f"{tag} b={b} t_start {ts[b]} vs {t0}"
        assert st[b, :Lb].tolist() == ss.tolist(), f"{tag} b={b} starts differ"
        assert en[b, :Lb].tolist() == se.tolist(), f"{tag} b={b} ends differ"
        np.testing.assert_allclose(sc[b, :Lb], ssc, rtol=SCORE_RTOL, atol=0, err_msg=f"{tag} b={b} scores")
    return n_ok


def test_align_golden_fixtures(golden_dir):
    """Every reference-generated case: trellis bit-exact (debug dump), path, spans, scores."""
    import ssak_b200
    for i, c in enumerate(load_cases(os.path.join(golden_dir, "align_golden.npz"))):
        e = torch.from_numpy(c["emission"])
        toks = c["tokens"].tolist()
        blank, fag = int(c["blank"]), bool(c["first_as_garbage"])
        T, L = e.shape[0], len(toks)
        tk = torch.tensor([toks], dtype=torch.int32).reshape(1, L)
        kw = {}
        if fag and L > 0:
            # column 0 of the garbage mode contains exp/log (:37); the fixture was produced by torch's CPU
            # kernels, so hand the kernel the CPU-computed column (what the HF / torchaudio back-ends of the
            # reference do: they move the emission to the CPU first).  The device-side default uses the same
            # torch ops on the GPU, like the reference would on a CUDA emission tensor.
            kw["col0"] = (1 - e[:, toks[0]].exp()).log().unsqueeze(0)
        res = _run(e.unsqueeze(0), tk, blank=blank, fag=fag, return_path=True, return_trellis=True, **kw)
        tr = res.trellis[0].cpu().numpy()
        assert np.array_equal(tr.view(np.int32), c["trellis"].view(np.int32)), f"case {i}: trellis bits"
        assert int(res.status[0]) == (0 if int(c["status"]) == 0 else 1), f"case {i}: status"
        assert int(res.t_start[0]) == int(c["t_start"]) or L == 0, f"case {i}: t_start"
        if int(c["status"]) != 0:
            continue
        assert res.starts[0].cpu().tolist() == c["seg_start"].tolist(), f"case {i}: starts"
        assert res.ends[0].cpu().tolist() == c["seg_end"].tolist(), f"case {i}: ends"
        np.testing.assert_allclose(res.scores[0].cpu().numpy(), c["seg_score"], rtol=SCORE_RTOL)
        ptok = res.path_token[0].cpu().numpy()
        frames = np.nonzero(ptok >= 0)[0]
        assert frames.tolist() == c["path_time"].tolist(), f"case {i}: path frames"
        assert ptok[frames].tolist() == c["path_token"].tolist(), f"case {i}: path tokens"
        np.testing.assert_allclose(res.path_prob[0].cpu().numpy()[frames], c["path_score"], rtol=SCORE_RTOL)
        # reference-shaped API
        trl = ssak_b200.get_trellis(e.cuda(), toks, blank_id=blank, first_as_garbage=fag)
        assert trl.size(0) == T + 1
        path = ssak_b200.backtrack(trl, e.cuda(), toks, blank_id=blank)
        assert [(p.token_index, p.time_index) for p in path] == list(zip(c["path_token"].tolist(), c["path_time"].tolist()))
        segs = ssak_b200.merge_repeats(list(range(L)), path)
        assert [(s.start, s.end) for s in segs] == list(zip(c["seg_start"].tolist(), c["seg_end"].tolist()))


@pytest.mark.parametrize("kind", ["random", "tie", "planted"])
@pytest.mark.parametrize("fag", [False, True])
def test_align_random_ragged_batches(kind, fag):
    from ssak_b200.synth import align_batch
    total = 0
    for seed, (B, T, V, Lmin, Lmax) in enumerate([(9, 60, 7, 1, 20), (6, 200, 50, 10, 70), (5, 97, 33, 1, 96),
                                                   (4, 300, 50, 100, 140)]):
        em, toks, el, tl = align_batch(B, T, V, Lmin, Lmax, 100 + seed, Tmin=max(1, T // 2), kind=kind)
        total += _check_batch(em, toks, el, tl, 0, fag, tag=f"{kind}/{fag}/{seed}")
    assert total > 0


def test_align_edge_cases():
    from ssak_b200.synth import align_batch
    # L > T, L == T, L == 1, L == 0 in one ragged batch; blank != 0
    em, toks, el, tl = align_batch(6, 40, 9, 1, 45, 7, kind="planted", blank=4)
    el = torch.tensor([40, 3, 40, 12, 40, 40], dtype=torch.int32)
    tl = torch.tensor([45, 5, 40, 12, 1, 0], dtype=torch.int32)
    res = _run(em, toks, el, tl, blank=4)
    assert res.status.cpu().tolist()[:2] == [1, 1] and int(res.status[5]) == 1
    _check_batch(em, toks, el, tl, blank=4, res=res, tag="edge")
    # SpeechBrain-style padded frames (-700, blank 0.0)
    em2, toks2, el2, tl2 = align_batch(3, 80, 12, 5, 20, 8, kind="planted")
    em2[:, -16:, :] = -700.0
    em2[:, -16:, 0] = 0.0
    assert _check_batch(em2, toks2, el2, tl2, tag="sb_padded") == 3


def test_align_non_contiguous_and_long_labels():
    from ssak_b200.synth import align_batch
    # emissions as a [T,B,V] -> [B,T,V] transposed view (strided batch/time axes)
    em, toks, el, tl = align_batch(4, 150, 50, 20, 60, 21, Tmin=100)
    em_tbv = em.transpose(0, 1).contiguous().cuda()
    import ssak_b200
    res = ssak_b200.forced_align(em_tbv.transpose(0, 1), toks, el, tl)
    _check_batch(em, toks, el, tl, res=res, tag="strided")
    # several states per lane and many warps (L+1 > 1024 -> K >= 4)
    em, toks, el, tl = align_batch(2, 2600, 50, 1100, 1300, 22, Tmin=2500)
    assert _check_batch(em, toks, el, tl, tag="longL") == 2
    # V = 1024 (BPE-sized rows, chunked ring)
    em, toks, el, tl = align_batch(3, 300, 1024, 40, 90, 23, Tmin=200)
    assert _check_batch(em, toks, el, tl, tag="V1024") == 3


def test_align_host_abi():
    """The host-buffer C entry point (what a non-torch caller binds)."""
    import ctypes as C
    import ssak_b200
    from ssak_b200.synth import align_batch
    L = ssak_b200.lib()
    em, toks, el, tl = align_batch(3, 90, 20, 5, 30, 31, Tmin=60)
    ctx = C.c_void_p()
    assert L.ssak_context_create(0, C.byref(ctx)) == 0
    B, T, V = em.shape
    Lm = toks.shape[1]
    emn, tkn = np.ascontiguousarray(em.numpy()), np.ascontiguousarray(toks.numpy().astype(np.int32))
    eln, tln = el.numpy().astype(np.int32), tl.numpy().astype(np.int32)
    st, en = np.zeros((B, Lm), np.int32), np.zeros((B, Lm), np.int32)
    sc, ts, status = np.zeros((B, Lm), np.float64), np.zeros(B, np.int32), np.zeros(B, np.int32)
    rc = L.ssak_forced_align_host(ctx, emn.ctypes.data, B, T, V, tkn.ctypes.data, Lm, eln.ctypes.data,
                                  tln.ctypes.data, 0, 0, None, st.ctypes.data, en.ctypes.data, sc.ctypes.data,
                                  ts.ctypes.data, status.ctypes.data)
    assert rc == 0
    L.ssak_context_destroy(ctx)
    for b in range(B):
        rcb, ss, se, ssc, t0 = O.align(emn[b, : eln[b]], tkn[b, : tln[b]].tolist())
        assert rcb == 0 and status[b] == 0 and ts[b] == t0
        assert st[b, : tln[b]].tolist() == ss.tolist() and en[b, : tln[b]].tolist() == se.tolist()


def test_align_full_size_properties():
    """C5-shaped batch (B=512 would take the oracle minutes): properties that hold at any size --
    spans are sorted, contiguous, inside [0, t_start], scores in (0, 1]; plus an oracle spot-check."""
    from ssak_b200.synth import align_batch
    em, toks, el, tl = align_batch(64, 750, 1024, 100, 200, 41, Tmin=600)
    res = _run(em, toks, el, tl)
    st, en, sc = res.starts.cpu().numpy(), res.ends.cpu().numpy(), res.scores.cpu().numpy()
    ts, status = res.t_start.cpu().numpy(), res.status.cpu().numpy()
    assert (status == 0).all()
    for b in range(em.shape[0]):
        Lb = int(tl[b])
        assert (st[b, :Lb] < en[b, :Lb]).all() and (en[b, : Lb - 1] == st[b, 1:Lb]).all()
        assert en[b, Lb - 1] == ts[b] <= int(el[b]) and st[b, 0] >= 0
        assert ((sc[b, :Lb] > 0) & (sc[b, :Lb] <= 1.0 + 1e-6)).all()
    idx = [0, 17, 63]
    _check_batch(em[idx], toks[idx], el[idx], tl[idx], res=None, tag="C5 spot")


def test_align_long_audio_cluster_sized_labels():
    """Config C3 shape at reduced duration: L ~ 8000 tokens (16 states per lane x 16 warps) and L ~ 3000
    (8 per lane), bit-exact against the oracle.  The full 10-minute T=30000 case is timed by tools/run_align.py."""
    from ssak_b200.synth import align_batch
    em, toks, el, tl = align_batch(2, 8600, 50, 7700, 8000, 51, Tmin=8400)
    assert _check_batch(em, toks, el, tl, tag="L8000") == 2
    em, toks, el, tl = align_batch(2, 4000, 50, 2800, 3100, 52, Tmin=3800, kind="tie")
    _check_batch(em, toks, el, tl, tag="L3000 tie")


@pytest.mark.parametrize("mode", ["barrier", "S1K4", "S2K4", "S4K4", "S8K4", "S3K8", "S1K8", "stages", "cluster", "S2K1",
                                  "S1K2", "S8K2", "S12K4", "default"])
def test_align_kernel_shapes_agree(mode, monkeypatch):
    """Every launch shape of the forward kernels gives the same bits: the per-frame-barrier kernel, and the
    wavefront kernel for 1..8 CTAs per utterance (cluster), 1, 2, 4 or 8 states per lane, minimal ring depth."""
    from ssak_b200.synth import align_batch
    env = {"barrier": {"SSAK_ALIGN_WAVE": "0"}, "S1K4": {"SSAK_ALIGN_S": "1", "SSAK_ALIGN_K": "4"},
           "S2K4": {"SSAK_ALIGN_S": "2", "SSAK_ALIGN_K": "4"}, "S4K4": {"SSAK_ALIGN_S": "4", "SSAK_ALIGN_K": "4"},
           "S8K4": {"SSAK_ALIGN_S": "8", "SSAK_ALIGN_K": "4"}, "S3K8": {"SSAK_ALIGN_S": "3", "SSAK_ALIGN_K": "8"},
           "S1K8": {"SSAK_ALIGN_S": "1", "SSAK_ALIGN_K": "8"},
           "stages": {"SSAK_ALIGN_S": "2", "SSAK_ALIGN_K": "4", "SSAK_ALIGN_STAGES": "5"},
           "cluster": {"SSAK_ALIGN_S": "4", "SSAK_ALIGN_K": "4", "SSAK_ALIGN_CLUSTER": "1"},
           "S2K1": {"SSAK_ALIGN_S": "2", "SSAK_ALIGN_K": "1"}, "S1K2": {"SSAK_ALIGN_S": "1", "SSAK_ALIGN_K": "2"},
           "S8K2": {"SSAK_ALIGN_S": "8", "SSAK_ALIGN_K": "2"}, "S12K4": {"SSAK_ALIGN_S": "12", "SSAK_ALIGN_K": "4"},
           "default": {}}[mode]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for seed, kind in ((61, "planted"), (62, "tie")):
        em, toks, el, tl = align_batch(3, 700, 50, 150, 520, seed, Tmin=500, kind=kind)
        tl[0] = 520
        tl[1] = 33          # most warps / CTAs of this utterance are idle
        _check_batch(em, toks, el, tl, tag=f"{mode}/{kind}")
    # T not a multiple of the chunk, T < chunk, first_as_garbage through the cross-CTA seam
    em, toks, el, tl = align_batch(4, 301, 20, 130, 290, 63, Tmin=5, kind="tie")
    el[0], tl[0] = 3, 2
    el[1], tl[1] = 301, 290
    _check_batch(em, toks, el, tl, fag=True, tag=f"{mode}/ragged")


@pytest.mark.parametrize("kind", ["random", "tie", "planted"])
@pytest.mark.parametrize("fag", [False, True])
def test_align_throughput_kernel(kind, fag, monkeypatch):
    """align_lane_kernel (one warp per utterance; the default from B = 2 x SMs on, forced here): 8 and 16 states per
    lane, V up to 128, ragged lengths, L > T, strided emissions -- bit-exact like the other kernels."""
    from ssak_b200.synth import align_batch
    monkeypatch.setenv("SSAK_ALIGN_LANE", "1")
    total = 0
    for seed, (B, T, V, Lmin, Lmax) in enumerate([(9, 60, 7, 1, 20), (6, 200, 50, 10, 70), (5, 97, 33, 1, 96),
                                                   (4, 300, 50, 100, 140), (3, 700, 128, 250, 255), (3, 900, 50, 256, 400),
                                                   (2, 1100, 97, 480, 511), (40, 90, 50, 5, 40)]):
        em, toks, el, tl = align_batch(B, T, V, Lmin, Lmax, 300 + seed, Tmin=max(1, T // 2), kind=kind)
        total += _check_batch(em, toks, el, tl, 0, fag, tag=f"lane/{kind}/{fag}/{seed}")
    assert total > 0
    if not fag:
        em, toks, el, tl = align_batch(6, 40, 9, 1, 45, 7, kind=kind, blank=4)
        el = torch.tensor([40, 3, 40, 12, 40, 40], dtype=torch.int32)
        tl = torch.tensor([45, 5, 40, 12, 1, 0], dtype=torch.int32)
        res = _run(em, toks, el, tl, blank=4)
        assert res.status.cpu().tolist()[:2] == [1, 1] and int(res.status[5]) == 1
        _check_batch(em, toks, el, tl, blank=4, res=res, tag="lane/edge")
        import ssak_b200
        em, toks, el, tl = align_batch(4, 150, 50, 20, 60, 21, Tmin=100, kind=kind)
        em_tbv = em.transpose(0, 1).contiguous().cuda()
        res = ssak_b200.forced_align(em_tbv.transpose(0, 1), toks, el, tl)
        _check_batch(em, toks, el, tl, res=res, tag="lane/strided")


def test_compute_alignment_adapter_and_word_positions():
    """The reference's entry point with its own signature (align_transcriptions.py:294-308) on an injected model front
    end, and tools/get_word_positions.py:14-43 on top of it: same segments as the from-emission variant and the oracle."""
    import ssak_b200
    from ssak_b200.synth import planted_emissions
    labels = ["<pad>", " "] + list("abcdefghijklmnopqrstuvwxyz'")
    dic = {c: i for i, c in enumerate(labels)}
    g = torch.Generator().manual_seed(5)
    transcript = "le chat dort"
    toks = [dic[c] for c in transcript]
    em = planted_emissions(90, len(labels), toks, g, 0)
    audio = torch.zeros(90 * 320)                      # 20 ms frames at 16 kHz
    model = object()
    calls = []

    def compute_log_probas(m, a):
        calls.append(("lp", m is model, len(a)))
        return em                                       # a CPU tensor, as the reference returns it

    def get_model_vocab(m):
        return labels, 0

    out = ssak_b200.compute_alignment(audio, transcript, model, compute_log_probas=compute_log_probas,
                                      get_model_vocab=get_model_vocab)
    lab, emission, trellis, segs, words = out
    assert lab == labels and emission.is_cuda and trellis.size(0) == 91 and calls == [("lp", True, 90 * 320)]
    rc, ss, se, sc, _ = O.align(em.numpy(), toks, 0, False)
    assert rc == 0 and [(x.start, x.end) for x in segs] == list(zip(ss.tolist(), se.tolist()))
    assert [w.label for w in words] == transcript.split()
    # transcript=None: decoded first (the reference prints it, :306-308)
    out2 = ssak_b200.compute_alignment(audio, None, model, compute_log_probas=compute_log_probas,
                                       get_model_vocab=get_model_vocab, decode_log_probas=lambda m, e: transcript)
    assert [(x.start, x.end) for x in out2[3]] == [(x.start, x.end) for x in segs]
    pos = list(ssak_b200.word_positions([audio], [transcript], model, 16000, compute_log_probas=compute_log_probas,
                                        get_model_vocab=get_model_vocab))
    ratio = len(audio) / (91 * 16000)
    assert [p["word"] for p in pos] == transcript.split()
    assert all(abs(p["start"] - w.start * ratio) < 1e-12 and abs(p["end"] - w.end * ratio) < 1e-12 and p["conf"] == w.score
               for p, w in zip(pos, words))


def test_windowed_long_audio_feeds_the_aligner():
    """a long recording goes through the acoustic model in windows (140 s = 7000 frames at 20 ms,
    ssak/infer/transformers_infer.py:259-265); the concatenated emissions -- seams included -- feed one long
    alignment (the C3 shape, shortened): same spans as aligning the whole emission, and as the CPU oracle."""
    import ssak_b200
    from ssak_b200.synth import align_batch
    em, toks, el, tl = align_batch(2, 17000, 50, 3800, 4200, 41, Tmin=16500)
    outs = []
    for b in range(2):
        Tb = int(el[b])

        def infer_sub(i, j, b=b):                       # 320 samples per frame
            return em[b : b + 1, i // 320 : j // 320].cuda()

        outs.append(ssak_b200.chunked_emission(infer_sub, Tb * 320, 7000 * 320)[0])
        assert outs[-1].shape[0] == Tb
    T = max(o.shape[0] for o in outs)
    batch = torch.zeros(2, T, 50, device="cuda")
    for b, o in enumerate(outs):
        batch[b, : o.shape[0]] = o
    res = ssak_b200.forced_align(batch, toks, el, tl)
    whole = ssak_b200.forced_align(em.cuda(), toks, el, tl)
    assert torch.equal(res.starts, whole.starts) and torch.equal(res.ends, whole.ends) and (res.status == 0).all()
    Tb, Lb = int(el[0]), int(tl[0])
    rc, ss, se, sc, t0 = O.align(em[0, :Tb].numpy(), toks[0, :Lb].tolist(), 0, False)
    assert rc == 0 and res.starts[0, :Lb].cpu().tolist() == ss.tolist() and res.ends[0, :Lb].cpu().tolist() == se.tolist()
    # greedy decode of the windowed emission (transformers_infer.py:84-85) equals the decode of the whole one
    ids, _ = ssak_b200.decode_chunked(lambda i, j: em[0:1, i // 320 : j // 320].cuda(), Tb * 320, 7000 * 320)
    assert ids == ssak_b200.ctc_greedy_decode(em[0:1, :Tb].cuda(), torch.ones(1), blank_id=0)[0]


def test_compute_alignments_batched_front_end():
    """compute_alignment from the emission onwards (align_transcriptions.py:310-402), batched: character ->
    token mapping with the loose fall-backs, sentinel character, word regrouping and score aggregation."""
    import ssak_b200
    labels = ["<pad>", "<s>", "</s>", "<unk>", " "] + list("abcdefghijklmnopqrstuvwxyz'-") + list("àâéèêëîïôùûç")
    V, blank = len(labels), 0
    dic = {c: i for i, c in enumerate(labels)}
    texts = ["bonjour à tous", ["c'est", "pas", "plus", "mal,", "euh"], "OUI Élise", ["allô", "allô"]]
    g = torch.Generator().manual_seed(77)
    ems, toks_all = [], []
    for tx in texts:
        chars = tx if isinstance(tx, str) else " ".join(tx)
        toks = [ssak_b200.loose_get_char_index(dic, c, dic[" "]) for c in chars]
        T = 3 * len(toks) + int(torch.randint(5, 40, (1,), generator=g))
        from ssak_b200.synth import planted_emissions
        ems.append(planted_emissions(T, V, toks, g, blank))
        toks_all.append(toks)
    out = ssak_b200.compute_alignments([e.cuda() for e in ems], texts, labels, blank)
    punct = set('!"#$%&()*+,./:;<=>?@[\\]^_`{|}~')
    for tx, e, toks, res in zip(texts, ems, toks_all, out):
        assert res is not None
        chars = tx if isinstance(tx, str) else " ".join(tx)
        rc, ss, se, sc, _ = O.align(e.numpy(), toks, blank, False)
        assert rc == 0
        char_segments, word_segments = res
        assert [(s.label, s.start, s.end) for s in char_segments] == [(chars[i], int(ss[i]), int(se[i])) for i in range(len(toks))]
        np.testing.assert_allclose([s.score for s in char_segments], sc, rtol=1e-6)
        # words: split on the space segments (:159-173) or by the given word list (:373-387)
        words = tx.split(" ") if isinstance(tx, str) else tx
        pos = 0
        assert [w.label for w in word_segments] == words
        for w, seg in zip(words, word_segments):
            idx = list(range(pos, pos + len(w)))
            keep = [i for i in idx if chars[i] != " " and chars[i] not in punct] if not isinstance(tx, str) else idx
            keep = keep or idx
            assert (seg.start, seg.end) == (int(ss[keep[0]]), int(se[keep[-1]]))
            lens = np.array([se[i] - ss[i] for i in keep], dtype=np.float64)
            assert abs(seg.score - float((sc[keep] * lens).sum() / lens.sum())) <= 1e-6
            pos += len(w) + 1
    # sentinel character on both ends (:336-339, :363-368) and the single-utterance wrapper
    lab2, em2, trellis, cs, ws = ssak_b200.compute_alignment_from_emission(ems[0].cuda(), texts[0], labels, blank,
                                                                              add_before_after=" ")
    assert "".join(s.label for s in cs) == texts[0] and trellis.size(0) == ems[0].shape[0] + 1
    with pytest.raises(RuntimeError, match="Failed to align"):
        ssak_b200.compute_alignment_from_emission(ems[0][:5].cuda(), texts[0], labels, blank)


def test_cut_kaldi_folder_batched_driver(tmp_path):
    """SURVEY 8 f-3 end to end: a Kaldi folder with short and long utterances is cut at word boundaries, several
    utterances per aligner launch; the output files equal what the oracle alignment of every utterance + the
    (reference-pinned) packing gives, in input order."""
    import ssak_b200
    from ssak_b200 import cutter
    from ssak_b200.synth import planted_emissions
    labels = ["<pad>", "<s>", "</s>", "<unk>", " "] + list("abcdefghijklmnopqrstuvwxyz'-") + list("àâéèêëîïôùûç")
    V, blank = len(labels), 0
    dic = {c: i for i, c in enumerate(labels)}
    texts = {"u1": "bonjour à tous", "u2": "oui", "u3": "c'est une phrase assez longue pour être coupée en deux ou trois",
             "u4": "et celle-ci aussi , vraiment très longue ! n'est-ce pas", "u5": "non"}
    durs = {"u1": 12.0, "u2": 1.0, "u3": 21.5, "u4": 17.25, "u5": 0.75}
    d = tmp_path / "in"
    d.mkdir()
    (d / "text").write_text("".join(f"{k} {v}\n" for k, v in texts.items()), encoding="utf-8")
    (d / "utt2spk").write_text("".join(f"{k} spk\n" for k in texts))
    (d / "utt2dur").write_text("".join(f"{k} {v}\n" for k, v in durs.items()))
    (d / "wav.scp").write_text("".join(f"{k} /data/{k}.wav\n" for k in texts))
    g = torch.Generator().manual_seed(91)
    ems = {}

    def emission_fn(uid, path, start, end):
        assert path == f"/data/{uid}.wav"
        words = cutter.regroup_isolated_punctuation(texts[uid].split())
        toks = [ssak_b200.loose_get_char_index(dic, c, dic[" "]) for c in " ".join(words)]
        T = int(round((end - start) * 50))                     # 20 ms frames
        ems[uid] = (planted_emissions(T, V, toks, g, blank), toks, words)
        return ems[uid][0].cuda(), end - start

    out = tmp_path / "out"
    stats = cutter.cut_kaldi_folder(str(d), str(out), emission_fn, labels, blank, max_duration=5.0, batch_size=2)
    assert stats["kept"] == 2 and stats["removed"] == 0 and stats["cut"] >= 6
    exp = {"text": "", "segments": "", "utt2dur": "", "utt2spk": ""}
    for uid in texts:
        if durs[uid] <= 5.0:
            exp["text"] += f"{uid} {texts[uid]}\n"
            exp["utt2spk"] += f"{uid} spk\n"
            exp["utt2dur"] += f"{uid} {durs[uid]}\n"
            exp["segments"] += f"{uid} {uid} 0 {durs[uid]}\n"
            continue
        em, toks, words = ems[uid]
        rc, ss, se, _, _ = O.align(em.numpy(), toks, blank, False)
        assert rc == 0
        spans, pos, chars = [], 0, " ".join(words)
        for w in words:   # a word spans its non-blank, non-punctuation characters (align_transcriptions.py:381-385)
            idx = list(range(pos, pos + len(w)))
            keep = [i for i in idx if chars[i] != " " and chars[i] not in cutter.PUNCTUATION] or idx
            spans.append((int(ss[keep[0]]), int(se[keep[-1]])))
            pos += len(w) + 1
        for c in cutter.pack_words(spans, words, em.shape[0], durs[uid] / em.shape[0], 0, 5.0):
            if c.written:
                nid = f"{uid}_cut{c.index:02}"
                exp["text"] += f"{nid} {c.transcript}\n"
                exp["utt2spk"] += f"{nid} spk\n"
                exp["utt2dur"] += f"{nid} {c.end - c.start:.3f}\n"
                exp["segments"] += f"{nid} {uid} {c.start:.3f} {c.end:.3f}\n"
    for k, v in exp.items():
        assert (out / k).read_text(encoding="utf-8") == v, k
