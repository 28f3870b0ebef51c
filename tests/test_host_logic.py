"""Host-side logic that needs no GPU: segment/word merging (vs the reference's own functions when
present), LPT sharding, argument checking, loud failure without CUDA."""
import numpy as np
import pytest
import torch

from oracle import ref_extract as R


def _random_path(rng, L, T):
    from ssak_b200 import Point
    onsets = np.sort(rng.choice(np.arange(T), size=L, replace=False))
    end = int(rng.integers(onsets[-1] + 1, T + 1))
    path = []
    for j in range(L):
        hi = onsets[j + 1] if j + 1 < L else end
        path += [Point(j, int(t), float(rng.random())) for t in range(onsets[j], hi)]
    return path


def test_merge_repeats_and_words_hand_case():
    from ssak_b200 import Point, merge_repeats, merge_words
    path = [Point(0, 2, 0.5), Point(0, 3, 1.0), Point(1, 4, 0.25), Point(2, 5, 0.5), Point(2, 6, 0.5), Point(3, 7, 1.0)]
    segs = merge_repeats("a bc", path)
    assert [(s.label, s.start, s.end) for s in segs] == [("a", 2, 4), (" ", 4, 5), ("b", 5, 7), ("c", 7, 8)]
    assert segs[0].score == 0.75 and segs[0].length == 2
    words = merge_words(segs)
    assert [(w.label, w.start, w.end) for w in words] == [("a", 2, 4), ("bc", 5, 8)]
    assert abs(words[1].score - (0.5 * 2 + 1.0 * 1) / 3) < 1e-12


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_merge_functions_match_reference():
    from ssak_b200 import merge_repeats, merge_words
    ns = R.namespace()
    rng = np.random.default_rng(0)
    for _ in range(50):
        L, T = int(rng.integers(1, 25)), int(rng.integers(25, 90))
        transcript = "".join(rng.choice(list("ab c'd "), size=L))
        path = _random_path(rng, L, T)
        ref_path = [ns["Point"](p.token_index, p.time_index, p.score) for p in path]
        ours, ref = merge_repeats(transcript, path), ns["merge_repeats"](transcript, ref_path)
        assert [(s.label, s.start, s.end, s.score) for s in ours] == [(s.label, s.start, s.end, s.score) for s in ref]
        ow, rw = merge_words(ours), ns["merge_words"](ref)
        assert [(s.label, s.start, s.end) for s in ow] == [(s.label, s.start, s.end) for s in rw]
        np.testing.assert_allclose([s.score for s in ow], [s.score for s in rw], rtol=1e-12)


def test_lpt_partition_balances_and_is_deterministic():
    from ssak_b200.shard import lattice_cost, length_buckets, lpt_partition
    rng = np.random.default_rng(1)
    il = rng.integers(300, 1501, size=256)
    tl = (0.27 * il).astype(int)
    costs = lattice_cost(il, tl)
    for ws in (2, 4, 8):
        parts = lpt_partition(costs, ws)
        assert sorted(i for p in parts for i in p) == list(range(256))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) / (sum(loads) / ws) < 1.02
        assert parts == lpt_partition(costs, ws)
    b = length_buckets(costs, 4)
    assert sorted(i for p in b for i in p) == list(range(256))
    assert max(costs[i] for i in b[0]) <= min(costs[i] for i in b[1])
    assert lattice_cost([10], [3], "align") == [40]


def test_product_fails_loudly_without_cuda_tensors():
    import ssak_b200
    with pytest.raises(RuntimeError, match="CUDA"):
        ssak_b200.forced_align(torch.zeros(1, 4, 3), torch.zeros(1, 2, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="CUDA"):
        ssak_b200.ctc_greedy_decode(torch.zeros(1, 4, 3), torch.ones(1))
    with pytest.raises(RuntimeError, match="CUDA"):
        ssak_b200.ctc_loss(torch.zeros(4, 1, 3), torch.zeros(1, 1, dtype=torch.long), [4], [1])


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_loose_get_char_index_matches_reference(capsys):
    from ssak_b200 import loose_get_char_index
    ref = R.namespace()["loose_get_char_index"]
    # The reference tries its fall-backs in the iteration order of a Python set (:410-414), which is arbitrary
    # when several fall-backs hit different labels; use a vocabulary where at most one of them can match.
    labels = ["<pad>", "|", " "] + list("abcdefghijklmnopqrstuvwxyz'")
    dictionary = {c: i for i, c in enumerate(labels)}
    for c in "aAzZéÉèÈçÇ'-!0 ñßœ|E":
        assert loose_get_char_index(dictionary, c, 2) == ref(dictionary, c, 2), c
    capsys.readouterr()
