"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference and torch CPU):
    python tests/golden/make_golden.py

  align_golden.npz  -- outputs of the reference's own get_trellis/backtrack/merge_repeats
                       (ssak/utils/align_transcriptions.py:27-157, executed unmodified through
                       oracle/ref_extract.py) on seeded synthetic emissions: random, tie-heavy
                       quantised, planted-alignment, first_as_garbage, blank != 0, L=1, L=T, L>T,
                       L=0, repeated characters, SpeechBrain-style -700 padded frames.
  ctc_golden.npz    -- torch.nn.functional.ctc_loss (CPU, fp32 and fp64) loss / per-sample nll /
                       gradient on seeded inputs: padded + flat targets, input_length < T,
                       target_length 0, infeasible utterances with zero_infinity, blank != 0,
                       forced repeats, all three reductions.
  greedy_golden.npz -- torch.argmax + itertools.groupby + drop blank (general.py:112,118).

The reference's own tests hold no tensor-level vectors for this path (all
goldens need hub checkpoints), so these reference-generated fixtures are the pin.
"""
from __future__ import annotations

import itertools
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_extract as R  # noqa: E402


def emissions(kind, T, V, tokens, g, blank=0):
    L = len(tokens)
    if kind == "tie":
        # exact binary fractions, deliberately NOT normalised -> ~9 % exactly tied cells
        return ((torch.round(2 * torch.randn(T, V, generator=g) * 2) / 2) - 8).numpy().astype(np.float32)
    lg = torch.randn(T, V, generator=g)
    if kind in ("planted", "sb_padded"):
        lg[:, blank] += 6
        if 0 < L <= T:
            on = torch.sort(torch.randperm(T, generator=g)[:L]).values
            lg[on, blank] -= 6
            lg[on, torch.tensor(tokens, dtype=torch.long)] += 6
    e = lg.log_softmax(-1)
    if kind == "sb_padded":
        # speechbrain_infer.py:237-242: padded frames set to -700 with blank at 0
        n = max(1, T // 5)
        e[-n:, :] = -700.0
        e[-n:, blank] = 0.0
    return e.numpy().astype(np.float32)


def make_align():
    rng = np.random.default_rng(20251018)
    cases = []
    spec = []
    for kind in ("random", "tie", "planted", "sb_padded"):
        for (T, V, L) in ((1, 3, 1), (7, 4, 3), (12, 5, 12), (40, 6, 9), (64, 50, 20), (33, 8, 32),
                          (150, 50, 40), (97, 11, 1)):
            spec.append((kind, T, V, L, 0, False))
    spec += [("random", 3, 4, 5, 0, False), ("planted", 10, 4, 0, 0, False),      # L>T, L=0 -> fail
             ("random", 30, 6, 8, 5, False), ("tie", 30, 6, 8, 2, False),         # blank != 0
             ("random", 8, 5, 3, 0, True), ("planted", 60, 9, 12, 0, True),
             ("tie", 50, 7, 10, 0, True), ("random", 45, 6, 9, 3, True)]          # first_as_garbage
    for i, (kind, T, V, L, blank, fag) in enumerate(spec):
        g = torch.Generator().manual_seed(1000 + i)
        toks = rng.integers(0, V, size=L)
        if L >= 4 and i % 2 == 0:
            toks[1] = toks[0]; toks[3] = toks[2]                                   # doubled chars
        e = emissions(kind, T, V, toks.tolist(), g, blank)
        ref = R.align(e, toks.tolist(), blank, fag, want_trellis=True)
        cases.append(dict(kind=kind, emission=e, tokens=toks.astype(np.int32), blank=blank,
                          first_as_garbage=fag, status=ref["status"], t_start=ref["t_start"],
                          trellis=ref["trellis"],
                          path_token=np.array([p[0] for p in ref["path"]], np.int32),
                          path_time=np.array([p[1] for p in ref["path"]], np.int32),
                          path_score=np.array([p[2] for p in ref["path"]], np.float64),
                          seg_start=np.array([s[1] for s in ref["segments"]], np.int32),
                          seg_end=np.array([s[2] for s in ref["segments"]], np.int32),
                          seg_score=np.array([s[3] for s in ref["segments"]], np.float64)))
    flat = {"n": np.array(len(cases))}
    for i, c in enumerate(cases):
        for k, v in c.items():
            flat[f"c{i}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "align_golden.npz"), **flat)
    print("align cases:", len(cases), "failures:", sum(c["status"] != 0 for c in cases))


def make_ctc():
    rng = np.random.default_rng(77)
    flat = {}
    n = 0
    for trial in range(14):
        T = int(rng.integers(2, 48)); B = int(rng.integers(1, 5)); V = int(rng.integers(3, 12))
        blank = int(rng.integers(0, V)) if trial % 3 == 1 else 0
        il = rng.integers(1, T + 1, size=B); il[0] = T
        tl = np.array([int(rng.integers(0, min(int(il[b]), 12) + 1)) for b in range(B)])
        if trial == 5:
            tl[:] = np.minimum(il, 12); il[-1] = max(1, tl[-1] - 1)                 # infeasible sample
        Smax = max(int(tl.max()), 1)
        labels = [c for c in range(V) if c != blank]
        tg = rng.choice(labels, size=(B, Smax))
        if trial % 4 == 0 and Smax >= 2:
            tg[:, 1] = tg[:, 0]
        g = torch.Generator().manual_seed(500 + trial)
        lp = torch.randn(T, B, V, generator=g).log_softmax(-1)
        zi = trial % 2 == 0 or trial == 5
        for red in ("none", "mean", "sum"):
            rec = dict(log_probs=lp.numpy(), targets=tg.astype(np.int64), input_lengths=il.astype(np.int64),
                       target_lengths=tl.astype(np.int64), blank=blank, zero_infinity=zi, reduction=red)
            for dt, key in ((torch.float32, "f32"), (torch.float64, "f64")):
                x = lp.to(dt).clone().requires_grad_(True)
                loss = F.ctc_loss(x, torch.tensor(tg), torch.tensor(il), torch.tensor(tl), blank=blank,
                                  reduction=red, zero_infinity=zi)
                loss.sum().backward()
                rec[f"loss_{key}"] = loss.detach().numpy()
                rec[f"grad_{key}"] = x.grad.numpy()
            for k, v in rec.items():
                flat[f"c{n}_{k}"] = np.asarray(v)
            n += 1
    flat["n"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "ctc_golden.npz"), **flat)
    print("ctc cases:", n)


def make_greedy():
    flat = {}
    n = 0
    for i, (B, T, V, blank, quant) in enumerate(((3, 20, 5, 0, False), (2, 50, 50, 0, False),
                                                  (4, 33, 7, 6, True), (1, 9, 1030, 0, False))):
        g = torch.Generator().manual_seed(900 + i)
        p = torch.randn(B, T, V, generator=g)
        if quant:
            p = torch.round(p)                                 # argmax ties -> first index
        p = p.log_softmax(-1) if not quant else p
        rel = torch.rand(B, generator=g) * 0.6 + 0.4
        rel[0] = 1.0
        outs = []
        for b in range(B):
            nfr = int(torch.round(rel[b] * T).item())
            ids = torch.argmax(p[b, :nfr], dim=-1).tolist()
            outs.append([k for k, _ in itertools.groupby(ids) if k != blank])
        flat[f"c{n}_probs"] = p.numpy()
        flat[f"c{n}_rel_lens"] = rel.numpy()
        flat[f"c{n}_blank"] = np.array(blank)
        flat[f"c{n}_argmax"] = torch.argmax(p, dim=-1).numpy().astype(np.int32)
        flat[f"c{n}_out_lens"] = np.array([len(o) for o in outs], np.int32)
        flat[f"c{n}_out"] = np.array([o + [-1] * (T - len(o)) for o in outs], np.int32)
        n += 1
    flat["n"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "greedy_golden.npz"), **flat)
    print("greedy cases:", n)


if __name__ == "__main__":
    assert R.available(), "needs /root/reference (build container only)"
    torch.set_num_threads(1)
    make_align()
    make_ctc()
    make_greedy()
