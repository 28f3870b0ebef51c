"""Seeded synthetic inputs for the CTC lattice path.  Generated on the CPU
so that the CPU checker and the kernels see identical bits."""
from __future__ import annotations

import torch


def planted_emissions(T: int, V: int, tokens, g: torch.Generator, blank: int = 0, boost: float = 6.0,
                      normalize: bool = True) -> torch.Tensor:
    """[T,V] log-probabilities with a planted monotone alignment of `tokens` (peaky, blank-dominated,
    like a trained CTC model)."""
    tokens = torch.as_tensor(tokens, dtype=torch.long)
    L = tokens.numel()
    lg = torch.randn(T, V, generator=g)
    lg[:, blank] += boost
    if 0 < L <= T:
        onset = torch.sort(torch.randperm(T, generator=g)[:L]).values
        lg[onset, blank] -= boost
        lg[onset, tokens] += boost
    return lg.log_softmax(-1) if normalize else lg


def tie_emissions(T: int, V: int, g: torch.Generator) -> torch.Tensor:
    """Exact binary fractions (not normalised): ~9 % of the trellis cells tie exactly."""
    return (torch.round(2 * torch.randn(T, V, generator=g) * 2) / 2) - 8


def ctc_batch(B: int, T: int, V: int, Lmin: int, Lmax: int, seed: int, Tmin: int | None = None,
              blank: int = 0, planted: bool = True):
    """-> log_probs [T,B,V] fp32, targets [B,Lmax] int64 (padded with blank... never read),
    input_lengths [B], target_lengths [B] (all CPU)."""
    g = torch.Generator().manual_seed(seed)
    Tmin = T if Tmin is None else Tmin
    tl = torch.randint(Lmin, Lmax + 1, (B,), generator=g)
    il = torch.randint(Tmin, T + 1, (B,), generator=g)
    il = torch.maximum(il, torch.minimum(2 * tl + 1, torch.tensor(T)))
    labels = torch.randint(0, V - 1, (B, max(Lmax, 1)), generator=g)
    labels = labels + (labels >= blank).long()  # uniform over the non-blank labels
    lp = torch.empty(T, B, V)
    for b in range(B):
        Tb, Lb = int(il[b]), int(tl[b])
        if planted:
            e = torch.randn(T, V, generator=g)
            e[:Tb] = planted_emissions(Tb, V, labels[b, :Lb], g, blank, normalize=False)
            lp[:, b] = e.log_softmax(-1)
        else:
            lp[:, b] = torch.randn(T, V, generator=g).log_softmax(-1)
    return lp, labels, il, tl


def align_batch(B: int, T: int, V: int, Lmin: int, Lmax: int, seed: int, Tmin: int | None = None,
                blank: int = 0, kind: str = "planted"):
    """-> emissions [B,T,V] fp32, tokens [B,Lmax] int32, emission_lengths [B], token_lengths [B]."""
    g = torch.Generator().manual_seed(seed)
    Tmin = T if Tmin is None else Tmin
    tl = torch.randint(Lmin, Lmax + 1, (B,), generator=g)
    el = torch.randint(Tmin, T + 1, (B,), generator=g)
    toks = torch.randint(0, V, (B, max(Lmax, 1)), generator=g).to(torch.int32)
    em = torch.empty(B, T, V)
    for b in range(B):
        Tb, Lb = int(el[b]), int(tl[b])
        if kind == "tie":
            em[b] = tie_emissions(T, V, g)
        elif kind == "random":
            em[b] = torch.randn(T, V, generator=g).log_softmax(-1)
        else:
            e = torch.randn(T, V, generator=g)
            e[:Tb] = planted_emissions(Tb, V, toks[b, :Lb].long(), g, blank, normalize=False)
            em[b] = e.log_softmax(-1)
    return em, toks[:, :Lmax], el.to(torch.int32), tl.to(torch.int32)
