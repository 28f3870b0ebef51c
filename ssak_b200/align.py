"""Forced alignment on the sm_100a kernels behind the reference's call signatures.

What it replaces (paths relative to the reference repository):
  * get_trellis / backtrack / merge_repeats / merge_words, Point, Segment --
    ssak/utils/align_transcriptions.py:27-70, 79-123, 141-157, 159-173, 72-76, 126-138
  * the emission-to-segments part of compute_alignment -- :310-402 (everything after the
    acoustic model has produced `emission`), as called from
    tools/align_audio_transcript.py:335 and tools/get_word_positions.py:33

`forced_align` is the batched entry (many utterances per launch); `get_trellis`/`backtrack`
keep the reference's one-utterance signatures.  The (T+1)x(L+1) trellis is never written to
HBM: `get_trellis` returns a stand-in that carries the alignment and answers `.size(0)`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib


@dataclass
class Point:  # align_transcriptions.py:72-76
    token_index: int
    time_index: int
    score: float


@dataclass
class Segment:  # align_transcriptions.py:126-138
    label: str
    start: int
    end: int
    score: float

    def __repr__(self):
        return f"{self.label}\t({self.score:4.2f}): [{self.start:5d}, {self.end:5d})"

    @property
    def length(self):
        return self.end - self.start


@dataclass
class AlignResult:
    """Batched output of `forced_align` (device tensors)."""
    starts: torch.Tensor      # int32 [B,Lmax]  Segment.start per token (-1: padding / failed)
    ends: torch.Tensor        # int32 [B,Lmax]  Segment.end
    scores: torch.Tensor      # float64 [B,Lmax] Segment.score
    status: torch.Tensor      # int32 [B]  0 ok, 1 "Failed to align"
    t_start: torch.Tensor     # int32 [B]  argmax_t trellis[t, L]
    path_token: Optional[torch.Tensor] = None   # int32 [B,Tmax]  Point.token_index per frame or -1
    path_prob: Optional[torch.Tensor] = None    # float32 [B,Tmax] Point.score per frame
    trellis: Optional[torch.Tensor] = None      # float32 [B,Tmax+1,Lmax+1] (debug only)


def forced_align(emissions, tokens, emission_lengths=None, token_lengths=None, blank_id=0,
                 first_as_garbage=False, return_path=False, return_trellis=False, col0=None) -> AlignResult:
    """Align B utterances in one launch.

    emissions [B,Tmax,V] fp32 CUDA log-probabilities, tokens [B,Lmax] integer ids, lengths [B]
    (default: full).  Semantics per utterance = reference get_trellis + backtrack + merge_repeats."""
    _lib.require_cuda(emissions, "emissions")
    if emissions.dim() != 3:
        raise RuntimeError("emissions must be [B, Tmax, V]")
    if emissions.dtype != torch.float32:
        emissions = emissions.float()
    if emissions.stride(2) != 1 and emissions.size(2) > 1:
        emissions = emissions.contiguous()
    B, Tmax, V = emissions.shape
    dev = emissions.device
    tokens = torch.as_tensor(tokens)
    if tokens.dim() != 2 or tokens.size(0) != B:
        raise RuntimeError("tokens must be [B, Lmax]")
    Lmax = tokens.size(1)
    tok = tokens.to(device=dev, dtype=torch.int32).contiguous()
    if tok.numel() == 0:
        tok = torch.zeros((B, 1), dtype=torch.int32, device=dev)
    def _len(x, full):
        if x is None:
            return torch.full((B,), full, dtype=torch.int32, device=dev)
        return torch.as_tensor(x).to(device=dev, dtype=torch.int32).contiguous()
    em_len, tok_len = _len(emission_lengths, Tmax), _len(token_lengths, Lmax)
    if not (0 <= int(blank_id) < V):
        raise RuntimeError("blank_id must be in the vocabulary range")
    L = _lib.lib()
    c0 = None
    if first_as_garbage:
        if col0 is not None:
            c0 = col0.to(device=dev, dtype=torch.float32).contiguous()
        elif Lmax > 0:
            # :37 with the same torch ops the reference applies to its emission tensor
            first = tok[:, :1].to(torch.int64).clamp_(0, V - 1)
            e0 = emissions.gather(2, first.view(B, 1, 1).expand(B, Tmax, 1)).squeeze(2)
            c0 = (1 - e0.exp()).log().contiguous()
        else:
            c0 = torch.zeros((B, max(Tmax, 1)), dtype=torch.float32, device=dev)
    ws_bytes = L.ssak_align_workspace_bytes(B, Tmax, Lmax)
    if ws_bytes == 0:
        raise _lib.SsakB200Error(f"forced_align: shape not supported (Tmax={Tmax}, Lmax={Lmax})")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    n = max(Lmax, 1)
    starts = torch.empty((B, n), dtype=torch.int32, device=dev)
    ends = torch.empty((B, n), dtype=torch.int32, device=dev)
    scores = torch.empty((B, n), dtype=torch.float64, device=dev)
    t_start = torch.empty(B, dtype=torch.int32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    ptok = torch.empty((B, Tmax), dtype=torch.int32, device=dev) if return_path else None
    pprob = torch.empty((B, Tmax), dtype=torch.float32, device=dev) if return_path else None
    dump = torch.empty((B, Tmax + 1, Lmax + 1), dtype=torch.float32, device=dev) if return_trellis else None
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        rc = L.ssak_forced_align(emissions.data_ptr(), B, Tmax, V, emissions.stride(0), emissions.stride(1),
                                 tok.data_ptr(), tok.stride(0), Lmax, em_len.data_ptr(), tok_len.data_ptr(),
                                 int(blank_id), int(bool(first_as_garbage)), _lib.ptr(c0), starts.data_ptr(),
                                 ends.data_ptr(), scores.data_ptr(), t_start.data_ptr(), status.data_ptr(),
                                 _lib.ptr(dump), _lib.ptr(ptok), _lib.ptr(pprob), ws.data_ptr(), ws_bytes,
                                 stream)
    _lib.check(rc, "ssak_forced_align")
    return AlignResult(starts[:, :Lmax], ends[:, :Lmax], scores[:, :Lmax], status, t_start, ptok, pprob, dump)


# ----------------------------------------------------------------- reference-shaped API (B = 1)
class Trellis:
    """What `get_trellis` returns instead of the [(T+1),(L+1)] tensor: the finished alignment.
    Supports the accesses the reference's callers make (`.size(0)`, `.shape`;
    tools/get_word_positions.py:34)."""

    def __init__(self, num_frames: int, num_tokens: int, result: AlignResult, dense=None):
        self.shape = (num_frames + 1, num_tokens + 1)
        self.result = result
        self.dense = dense  # the real tensor when get_trellis(..., materialize=True)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]


def get_trellis(emission, tokens: Sequence[int], blank_id=0, first_as_garbage=False, materialize=False):
    """align_transcriptions.py:27-70 signature.  emission [T,V] CUDA tensor, tokens list[int]."""
    _lib.require_cuda(emission, "emission")
    toks = torch.tensor([list(tokens)], dtype=torch.int32).reshape(1, -1)
    res = forced_align(emission.unsqueeze(0), toks, blank_id=blank_id, first_as_garbage=first_as_garbage,
                       return_path=True, return_trellis=materialize)
    return Trellis(emission.size(0), toks.size(1), res, res.trellis[0] if materialize else None)


def backtrack(trellis: Trellis, emission=None, tokens=None, blank_id=0) -> List[Point]:
    """align_transcriptions.py:79-123 signature: the path as Points, or the reference's
    RuntimeError when the alignment failed."""
    res = trellis.result
    if int(res.status[0].item()) != 0:
        raise RuntimeError("Failed to align (not enough tokens for the duration?)")
    ptok = res.path_token[0].cpu()
    pprob = res.path_prob[0].cpu()
    idx = torch.nonzero(ptok >= 0).flatten().tolist()
    return [Point(int(ptok[f]), f, float(pprob[f])) for f in idx]


def merge_repeats(transcript, path: List[Point]) -> List[Segment]:
    """align_transcriptions.py:141-157: run-length merge of the path into per-token segments."""
    segments: List[Segment] = []
    run: List[Point] = []
    for p in path + [None]:
        if run and (p is None or p.token_index != run[0].token_index):
            segments.append(Segment(transcript[run[0].token_index], run[0].time_index,
                                    run[-1].time_index + 1, sum(q.score for q in run) / len(run)))
            run = []
        if p is not None:
            run.append(p)
    return segments


def _word_from(segs: List[Segment], label: Optional[str] = None) -> Segment:
    total = sum(s.length for s in segs)
    score = sum(s.score * s.length for s in segs) / total
    return Segment("".join(s.label for s in segs) if label is None else label, segs[0].start, segs[-1].end, score)


def merge_words(segments: List[Segment], separator=" ") -> List[Segment]:
    """align_transcriptions.py:159-173: group character segments into words at `separator`."""
    words, cur = [], []
    for seg in segments:
        if seg.label == separator:
            if cur:
                words.append(_word_from(cur))
            cur = []
        else:
            cur.append(seg)
    if cur:
        words.append(_word_from(cur))
    return words


# Characters that do not count towards a word's score (ssak/utils/text_basic.py:15-16:
# ASCII punctuation plus CJK / Arabic / typographic marks, keeping "-" and "'").
_PUNCT_EXTRA = ("\u3002\uff0c\uff01\uff1f\uff1a\u201d\u3001\u2026"   # CJK stops, closing quote, ellipsis
                "\u061f\u060c\u061b"                                      # Arabic ? , ;
                "\u2014\u00ab\u00b0\u00bb\u00d7\u2039\u203a\u2022\u201c\u2013\u2018\u2033")


def _punctuation() -> frozenset:
    import string
    return frozenset(string.punctuation + _PUNCT_EXTRA) - {"-", "'"}


def segments_from_result(res: AlignResult, b: int, transcript: str) -> List[Segment]:
    """Character segments of utterance b (what merge_repeats returns in the reference)."""
    n = len(transcript)
    st, en, sc = res.starts[b, :n].tolist(), res.ends[b, :n].tolist(), res.scores[b, :n].tolist()
    return [Segment(transcript[i], st[i], en[i], sc[i]) for i in range(n)]


# --------------------------------------------------------------------------------------------------
# compute_alignment (align_transcriptions.py:294-402) from the emission onwards, one utterance or a batch
# (SURVEY.md section 8 f-2).  The acoustic-model front end (emission, labels, blank_id) stays with the
# reference: ssak.infer.general.compute_log_probas / get_model_vocab.
# --------------------------------------------------------------------------------------------------
def _ascii_fold(c: str) -> str:
    """Closest ASCII equivalent of a character (ssak/utils/text_basic.py:191-196)."""
    import unicodedata
    return unicodedata.normalize("NFKD", c).encode("ascii", "ignore").decode("ascii")


_MISSING_LABELS = set()


def loose_get_char_index(dictionary, c, default):
    """Vocabulary index of character `c` with the reference's fall-backs (:406-423): exact, lower / upper
    case, ASCII-folded (and its cases); otherwise `default` (warned once per character)."""
    i = dictionary.get(c)
    if i is not None:
        return i
    folded = _ascii_fold(c)
    for c2 in (c.lower(), c.upper(), folded, folded.lower(), folded.upper()):
        i = dictionary.get(c2)
        if i is not None:
            return i
    if c not in _MISSING_LABELS:
        _MISSING_LABELS.add(c)
        print("WARNING: cannot find label " + c)
    return default


def _prepare_transcript(transcript, labels, blank_id, add_before_after):
    if isinstance(transcript, str):
        chars, words = transcript, None
    else:
        assert isinstance(transcript, list), f"Got unexpected transcript (of type {type(transcript)})"
        for w in transcript:
            assert isinstance(w, str), f"Got unexpected type {type(w)} (not a string)"
        chars, words = " ".join(transcript), transcript
    space_id = labels.index(" ") if " " in labels else blank_id          # :332-334
    if add_before_after:                                                 # :336-339
        assert len(add_before_after) == 1 and add_before_after in labels
        chars = add_before_after + chars + add_before_after
    dictionary = {c: i for i, c in enumerate(labels)}
    tokens = [loose_get_char_index(dictionary, c, space_id) for c in chars]
    return chars, words, [t for t in tokens if t is not None]


def _segments_to_words(char_segments, chars, words, add_before_after):
    """:363-387: strip the sentinel characters, then group characters into words."""
    if add_before_after:
        assert char_segments[0].label == add_before_after and char_segments[-1].label == add_before_after
        char_segments = char_segments[1:-1]
        chars = chars[1:-1]
    if words is None:
        return char_segments, merge_words(char_segments)
    punct = _punctuation()
    out, i2 = [], -1
    for word in words:
        i1 = i2 + 1
        i2 = i1 + len(word)
        segs1 = char_segments[i1:i2]
        assert "".join(s.label for s in segs1) == word
        segs = [s for s in segs1 if s.label != " " and s.label not in punct] or segs1   # :381-385
        out.append(_word_from(segs, word))
    return char_segments, out


def compute_alignments(emissions, transcripts, labels, blank_id, add_before_after=None, first_as_garbage=False):
    """Batched compute_alignment: B utterances, ONE launch.

    emissions: list of [T_i, V] CUDA tensors (or a padded [B,Tmax,V] tensor with `emission_lengths` given as a
    second element of a tuple); transcripts: list of str or list[str] (words).  Returns a list with, per
    utterance, (char_segments, word_segments) or None where the reference raises "Failed to align"."""
    lengths = None
    if isinstance(emissions, tuple):
        emissions, lengths = emissions
    if isinstance(emissions, (list, tuple)):
        lengths = [int(e.shape[0]) for e in emissions]
        Tmax, V = max(lengths), emissions[0].shape[1]
        em = emissions[0].new_zeros((len(emissions), Tmax, V))
        for i, e in enumerate(emissions):
            em[i, : e.shape[0]] = e
    else:
        em = emissions
        if lengths is None:
            lengths = [em.shape[1]] * em.shape[0]
    labels = list(labels)[: em.shape[2]]                                 # :341
    prepared = [_prepare_transcript(t, labels, blank_id, add_before_after) for t in transcripts]
    Lmax = max((len(p[2]) for p in prepared), default=0)
    toks = torch.zeros((len(prepared), max(Lmax, 1)), dtype=torch.int32)
    for i, p in enumerate(prepared):
        toks[i, : len(p[2])] = torch.tensor(p[2], dtype=torch.int32)
    res = forced_align(em, toks[:, :Lmax] if Lmax else toks[:, :0], torch.tensor(lengths, dtype=torch.int32),
                       torch.tensor([len(p[2]) for p in prepared], dtype=torch.int32), blank_id=blank_id,
                       first_as_garbage=first_as_garbage)
    status = res.status.cpu().tolist()
    st, en, sc = res.starts.cpu(), res.ends.cpu(), res.scores.cpu()
    out = []
    for i, (chars, words, tk) in enumerate(prepared):
        if status[i] != 0:
            out.append(None)
            continue
        n = len(tk)
        segs = [Segment(chars[j], int(st[i, j]), int(en[i, j]), float(sc[i, j])) for j in range(n)]
        out.append(_segments_to_words(segs, chars, words, add_before_after))
    return out


def compute_alignment_from_emission(emission, transcript, labels, blank_id, add_before_after=None,
                                    first_as_garbage=False):
    """compute_alignment (:294-402) after `emission = compute_log_probas(model, audio)`:
    -> (labels, emission, trellis, char_segments, word_segments); raises the reference's RuntimeError."""
    out = compute_alignments([emission], [transcript], labels, blank_id, add_before_after, first_as_garbage)[0]
    if out is None:
        raise RuntimeError("Failed to align (not enough tokens for the duration?)")
    char_segments, word_segments = out
    n_tok = len(char_segments) + (2 if add_before_after else 0)
    trellis = Trellis(emission.shape[0], n_tok - (2 if add_before_after else 0), None)   # :367 drops the sentinels
    return list(labels)[: emission.shape[1]], emission, trellis, char_segments, word_segments


def compute_alignment(audio, transcript, model, add_before_after=None, first_as_garbage=False, plot=False, verbose=False,
                      *, compute_log_probas=None, get_model_vocab=None, decode_log_probas=None):
    """The reference's entry point, same signature and return value (align_transcriptions.py:294-402; callers:
    tools/align_audio_transcript.py:335, tools/get_word_positions.py:33):
        labels, emission, trellis, char_segments, word_segments = compute_alignment(audio, transcript, model)
    The model-side front end stays with the reference -- `ssak.infer.general.compute_log_probas`,
    `get_model_vocab` and (for transcript=None) `decode_log_probas` are imported from the reference package when it is
    installed, or passed in as callables (any `model` object goes straight through to them).  Everything from the
    emission on runs on the GPU aligner.  `plot` is accepted for signature compatibility (plots stay with the
    reference: align_transcriptions.py:176-292)."""
    if compute_log_probas is None or get_model_vocab is None:
        try:
            from ssak.infer import general as _general     # the reference package
        except ImportError as err:
            raise ImportError("compute_alignment(audio, transcript, model) needs the reference's model front end: install "
                              "ssak, or pass compute_log_probas= and get_model_vocab= callables") from err
        compute_log_probas = compute_log_probas or _general.compute_log_probas
        get_model_vocab = get_model_vocab or _general.get_model_vocab
        decode_log_probas = decode_log_probas or _general.decode_log_probas
    emission = compute_log_probas(model, audio)                                          # :304
    if not emission.is_cuda:
        emission = emission.cuda()
    if transcript is None:                                                               # :306-308
        if decode_log_probas is None:
            raise ValueError("transcript=None needs decode_log_probas")
        transcript = decode_log_probas(model, emission)
        print("Transcript:", transcript)
    labels, blank_id = get_model_vocab(model)                                            # :330
    out = compute_alignment_from_emission(emission, transcript, labels, blank_id, add_before_after, first_as_garbage)
    if verbose:
        for w in out[4]:
            print(w)
    return out


def word_positions(audios, annotations, model, sample_rate, **front_end):
    """tools/get_word_positions.py:14-43: one dict per word with start / end in seconds and the confidence."""
    for audio, transcript in zip(audios, annotations):
        _, _, trellis, _, word_segments = compute_alignment(audio, transcript, model, **front_end)
        ratio = len(audio) / (trellis.size(0) * sample_rate)                             # :34 (T + 1 rows, like the reference)
        all_words = transcript.split()
        assert len(all_words) == len(word_segments)
        for word, segment in zip(all_words, word_segments):
            yield {"word": word, "start": segment.start * ratio, "end": segment.end * ratio, "conf": segment.score}
