"""Greedy CTC decode on the GPU vs the fixtures / oracle: bit-exact (integer outputs)."""
import itertools
import os

import numpy as np
import pytest
import torch

from conftest import load_cases
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_greedy_golden(golden_dir):
    import ssak_b200
    for c in load_cases(os.path.join(golden_dir, "greedy_golden.npz")):
        p = torch.from_numpy(c["probs"]).cuda()
        rel = torch.from_numpy(c["rel_lens"])
        blank = int(c["blank"])
        out = ssak_b200.ctc_greedy_decode(p, rel, blank_id=blank)
        for b in range(p.shape[0]):
            assert out[b] == c["out"][b, : c["out_lens"][b]].tolist()
        assert ssak_b200.argmax_ids(p).cpu().numpy().tolist() == c["argmax"].tolist()


@pytest.mark.parametrize("shape", [(5, 33, 50), (3, 40, 1024), (2, 17, 3), (4, 64, 1), (3, 50, 257), (2, 30, 36)])
def test_greedy_random(shape):
    import ssak_b200
    B, T, V = shape
    g = torch.Generator().manual_seed(sum(shape))
    p = torch.round(torch.randn(B, T, V, generator=g) * 2) / 2      # many exact ties -> first index wins
    n = torch.randint(0, T + 1, (B,), generator=g)
    blank = V - 1
    ids, out, lens = ssak_b200.greedy_ids(p.cuda(), n, blank)
    assert torch.equal(ids.cpu().long(), torch.argmax(p, -1))
    for b in range(B):
        exp, _ = O.greedy(p[b].numpy(), int(n[b]), blank)
        assert out[b, : int(lens[b])].cpu().tolist() == exp
        ref = [k for k, _ in itertools.groupby(torch.argmax(p[b, : int(n[b])], -1).tolist()) if k != blank]
        assert exp == ref
    # negative blank id counts from the end (SpeechBrain), strided input
    q = p.transpose(0, 1).contiguous().cuda().transpose(0, 1)
    dec = ssak_b200.ctc_greedy_decode(q, torch.ones(B), blank_id=-1)
    for b in range(B):
        assert dec[b] == [k for k, _ in itertools.groupby(torch.argmax(p[b], -1).tolist()) if k != V - 1]


def test_greedy_host_abi():
    import ctypes as C
    import ssak_b200
    L = ssak_b200.lib()
    g = torch.Generator().manual_seed(3)
    p = torch.randn(3, 25, 40, generator=g).numpy()
    ctx = C.c_void_p()
    assert L.ssak_context_create(0, C.byref(ctx)) == 0
    ids, out, lens = np.zeros((3, 25), np.int32), np.zeros((3, 25), np.int32), np.zeros(3, np.int32)
    assert L.ssak_ctc_greedy_host(ctx, p.ctypes.data, 3, 25, 40, None, 0, ids.ctypes.data, out.ctypes.data,
                                  lens.ctypes.data) == 0
    L.ssak_context_destroy(ctx)
    for b in range(3):
        exp, fid = O.greedy(p[b], 25, 0)
        assert out[b, : lens[b]].tolist() == exp and ids[b].tolist() == fid.tolist()
