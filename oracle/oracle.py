"""ctypes/numpy front-end of the CPU oracle (oracle/ssak_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from ssak_b200/.

Every function restates a piece of the reference (paths relative to /root/reference):
  get_trellis / backtrack / merge_repeats  <- ssak/utils/align_transcriptions.py:27-70,79-123,141-157
  ctc_loss                                <- torch.nn.functional.ctc_loss as configured at
                                             ssak/train/transformers/wav2vec_train.py:313-325
  greedy                                  <- ssak/infer/general.py:112,118
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libssak_oracle.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile oracle/ssak_oracle.c with gcc (make -C oracle)."""
    src = os.path.join(_HERE, "ssak_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_get_trellis.argtypes = [_f32p, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, _f32p]
        L.oracle_get_trellis.restype = None
        L.oracle_backtrack.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _i32p, C.c_int, C.c_int,
                                       _i32p, _i32p, _f32p, C.POINTER(C.c_int)]
        L.oracle_backtrack.restype = C.c_int
        L.oracle_merge_repeats.argtypes = [_i32p, _i32p, _f32p, C.c_int, _i32p, _i32p, _i32p, _f64p]
        L.oracle_merge_repeats.restype = C.c_int
        L.oracle_align.argtypes = [_f32p, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, _i32p, _i32p, _f64p, C.POINTER(C.c_int),
                                   C.POINTER(C.c_int)]
        L.oracle_align.restype = C.c_int
        for suf, rp in (("f32", _f32p), ("f64", _f64p)):
            fa = getattr(L, f"oracle_ctc_alpha_{suf}")
            fa.argtypes = [rp, C.c_int64, C.c_int, _i32p, C.c_int, C.c_int, rp]
            fa.restype = C.c_double
            fg = getattr(L, f"oracle_ctc_grad_{suf}")
            fg.argtypes = [rp, C.c_int64, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, rp,
                           C.c_double, C.c_double, C.c_int, rp, rp, C.c_int64]
            fg.restype = None
        L.oracle_greedy.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, _i32p]
        L.oracle_greedy.restype = C.c_int
        _lib = L
    return _lib


class AlignmentFailure(RuntimeError):
    """The reference's RuntimeError("Failed to align ...") (align_transcriptions.py:121-122)."""


def _tok(tokens) -> np.ndarray:
    t = np.ascontiguousarray(np.asarray(tokens, dtype=np.int32).reshape(-1))
    return t if t.size else np.zeros(1, np.int32)[:0].copy()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def garbage_col0(emission: np.ndarray, tok0: int) -> np.ndarray:
    """Column 0 of the trellis in first_as_garbage mode, computed with the reference's own
    torch CPU ops (align_transcriptions.py:37) so libm vs Sleef rounding is not an issue."""
    import torch
    e = torch.from_numpy(np.ascontiguousarray(emission))
    return (1 - e[:, tok0].exp()).log().numpy().astype(np.float32)


def get_trellis(emission, tokens, blank_id=0, first_as_garbage=False) -> np.ndarray:
    """align_transcriptions.py:27-70 -> [(T+1),(L+1)] fp32."""
    e = np.ascontiguousarray(emission, dtype=np.float32)
    T, V = e.shape
    tok = _tok(tokens)
    Ltok = len(tok)
    col0 = None
    if first_as_garbage and Ltok > 0:
        col0 = np.ascontiguousarray(garbage_col0(e, int(tok[0])))
    tr = np.empty((T + 1, Ltok + 1), np.float32)
    tokp = tok if Ltok else np.zeros(1, np.int32)
    lib().oracle_get_trellis(e, T, V, tokp, Ltok, int(blank_id), int(bool(first_as_garbage)),
                             _ptr(col0), tr)
    return tr


@dataclass
class Point:  # align_transcriptions.py:72-76
    token_index: int
    time_index: int
    score: float


@dataclass
class Segment:  # align_transcriptions.py:126-138 (label replaced by the token index)
    token_index: int
    start: int
    end: int
    score: float


def backtrack(trellis, emission, tokens, blank_id=0):
    """align_transcriptions.py:79-123 -> list[Point]; raises AlignmentFailure."""
    e = np.ascontiguousarray(emission, dtype=np.float32)
    tr = np.ascontiguousarray(trellis, dtype=np.float32)
    T, V = e.shape
    tok = _tok(tokens)
    Ltok = len(tok)
    cap = max(T, 1)
    pt, pi, ps = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float32)
    n = C.c_int(0)
    tokp = tok if Ltok else np.zeros(1, np.int32)
    rc = lib().oracle_backtrack(tr, e, T, V, tokp, Ltok, int(blank_id), pt, pi, ps, C.byref(n))
    if rc != 0:
        raise AlignmentFailure("Failed to align (not enough tokens for the duration?)")
    return [Point(int(pt[i]), int(pi[i]), float(ps[i])) for i in range(n.value)]


def merge_repeats(path):
    """align_transcriptions.py:141-157 (labels left as token indices)."""
    n = len(path)
    if n == 0:
        return []
    pt = np.array([p.token_index for p in path], np.int32)
    pi = np.array([p.time_index for p in path], np.int32)
    ps = np.array([p.score for p in path], np.float32)
    st, ss, se, sc = (np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32),
                      np.zeros(n, np.float64))
    k = lib().oracle_merge_repeats(pt, pi, ps, n, st, ss, se, sc)
    return [Segment(int(st[i]), int(ss[i]), int(se[i]), float(sc[i])) for i in range(k)]


def align(emission, tokens, blank_id=0, first_as_garbage=False):
    """get_trellis + backtrack + merge_repeats for one utterance.

    Returns (status, starts[L] int32, ends[L] int32, scores[L] float64, t_start) with status 0
    on success and -1 for the reference's "Failed to align" error (outputs then undefined)."""
    e = np.ascontiguousarray(emission, dtype=np.float32)
    T, V = e.shape
    tok = _tok(tokens)
    Ltok = len(tok)
    col0 = None
    if first_as_garbage and Ltok > 0:
        col0 = np.ascontiguousarray(garbage_col0(e, int(tok[0])))
    cap = max(T, 1)
    ss, se, sc = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float64)
    n, ts = C.c_int(0), C.c_int(0)
    tokp = tok if Ltok else np.zeros(1, np.int32)
    rc = lib().oracle_align(e, T, V, tokp, Ltok, int(blank_id), int(bool(first_as_garbage)),
                            _ptr(col0), ss, se, sc, C.byref(n), C.byref(ts))
    if rc == -2:
        raise MemoryError("oracle_align: trellis allocation failed")
    k = n.value
    return rc, ss[:k].copy(), se[:k].copy(), sc[:k].copy(), ts.value


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean",
             zero_infinity=False, dtype=np.float32, want_grad=True, grad_out=None):
    """torch.nn.functional.ctc_loss semantics on numpy arrays.

    log_probs [T,B,V]; targets [B,Smax] padded or 1-D concatenated; lengths [B].
    Returns (loss, nll[B] float64, grad[T,B,V] or None).  grad is d(loss)/d(logits feeding
    log_softmax) in torch's convention, scaled by the reduction and by
    grad_out (scalar, or [B] for reduction='none')."""
    lp = np.ascontiguousarray(log_probs, dtype=dtype)
    T, B, V = lp.shape
    il = np.asarray(input_lengths, np.int64).reshape(-1)
    tl = np.asarray(target_lengths, np.int64).reshape(-1)
    tg = np.asarray(targets)
    if tg.ndim == 2:
        tgts = [np.ascontiguousarray(tg[b, : tl[b]], np.int32) for b in range(B)]
    else:
        off = np.concatenate([[0], np.cumsum(tl)])
        tgts = [np.ascontiguousarray(tg[off[b]: off[b + 1]], np.int32) for b in range(B)]
    if not (0 <= blank < V):
        raise RuntimeError("blank must be in label range")
    if (il > T).any() or (il < 0).any():
        raise RuntimeError("Expected input_lengths to have value at most T")
    suf = "f32" if dtype == np.float32 else "f64"
    fa, fg = getattr(lib(), f"oracle_ctc_alpha_{suf}"), getattr(lib(), f"oracle_ctc_grad_{suf}")
    nll = np.zeros(B, np.float64)
    alphas = []
    for b in range(B):
        Lb, Tb = int(tl[b]), int(il[b])
        a = np.empty((max(Tb, 1), 2 * Lb + 1), dtype)
        tgt = tgts[b] if Lb else np.zeros(1, np.int32)
        # utterance b's row t starts at flat offset t*(B*V) + b*V
        nll[b] = fa(lp.reshape(-1)[b * V:], B * V, Tb, tgt, Lb, int(blank), a.reshape(-1))
        alphas.append(a)
    nll_out = nll.copy()
    inf_mask = np.isinf(nll_out)
    if zero_infinity:
        nll_out[inf_mask] = 0.0
    if reduction == "none":
        loss = nll_out.astype(dtype)
        gs = np.ones(B) if grad_out is None else np.broadcast_to(np.asarray(grad_out, np.float64), (B,))
    elif reduction == "sum":
        loss = dtype(nll_out.astype(dtype).sum())
        gs = np.full(B, 1.0 if grad_out is None else float(grad_out))
    elif reduction == "mean":
        loss = dtype((nll_out.astype(dtype) / np.maximum(tl, 1).astype(dtype)).mean())
        gs = (1.0 if grad_out is None else float(grad_out)) / (np.maximum(tl, 1) * B)
    else:
        raise ValueError(reduction)
    grad = None
    if want_grad:
        grad = np.zeros((T, B, V), dtype)
        gflat = grad.reshape(-1)
        for b in range(B):
            Lb, Tb = int(tl[b]), int(il[b])
            beta = np.empty((max(Tb, 1), 2 * Lb + 1), dtype)
            tgt = tgts[b] if Lb else np.zeros(1, np.int32)
            fg(lp.reshape(-1)[b * V:], B * V, T, Tb, V, tgt, Lb, int(blank), alphas[b].reshape(-1),
               float(nll[b]), float(gs[b]), int(bool(zero_infinity)), beta.reshape(-1),
               gflat[b * V:], B * V)
    return loss, nll, grad


def greedy(probs, n_frames=None, blank_id=0):
    """argmax + collapse repeats + drop blank for one utterance [T,V] -> (ids list, frame argmax)."""
    p = np.ascontiguousarray(probs, dtype=np.float32)
    T, V = p.shape
    n = T if n_frames is None else int(n_frames)
    fid = np.zeros(max(n, 1), np.int32)
    out = np.zeros(max(n, 1), np.int32)
    k = lib().oracle_greedy(p, V, n, V, int(blank_id), fid.ctypes.data_as(C.c_void_p), out)
    return out[:k].tolist(), fid[:n].copy()
