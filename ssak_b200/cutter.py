"""Kaldi-folder cutter around the batched aligner.

What it mirrors (paths relative to the reference repository):
  * the on-disk formats `text`, `utt2spk`, `utt2dur`, `segments`, `wav.scp` -- README.md:10-22, parser
    ssak/utils/kaldi.py:8-37, readers tools/align_audio_transcript.py:190-238
  * the word-packing step of tools/align_audio_transcript.py:383-435 (`add_segment` and its loop): aligned words
    are packed greedily into cuts of at most `max_duration` seconds, frame -> second conversion through
    `ratio = len(audio) / (num_frames * sample_rate)` (:371), first / last word stretched to the utterance when
    the time stamps are not refined (:399-401), isolated punctuation given zero length (:403-404)
  * the short-utterance pass-through (:301-307)
The acoustic model, audio loading and text normalisation stay with the reference: the driver takes a callable
that returns the emission of an utterance.  Unlike the reference loop (one utterance per aligner call, :335),
`cut_kaldi_folder` collects `batch_size` utterances per launch of the wavefront aligner.

Host logic only (pure Python; the alignment itself is ssak_b200.compute_alignments -> libssak_b200.so).
"""
from __future__ import annotations

import os
import re
import string
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

# ssak/utils/text_basic.py:15-16
_PUNCTUATION_STRONG = string.punctuation + "。，！？：”、…" + "؟،؛" + "—" + "«°»×‹›•“–‘″‘"
PUNCTUATION = "".join(c for c in _PUNCTUATION_STRONG if c not in ["-", "'"])


# ------------------------------------------------------------------------------------------- Kaldi folders
def parse_kaldi_wavscp(wavscp: str) -> Dict[str, str]:
    """wav.scp -> {wav id: path}: plain `id path`, quoted paths, `sox file ... |` and `flac ... file |` pipes,
    environment variables expanded (ssak/utils/kaldi.py:8-37)."""
    wav = {}
    with open(wavscp) as f:
        for line in f:
            fields = [x for x in line.strip().split() if x != "|"]
            if not fields:
                continue
            wavid = fields[0]
            if line.find("'") >= 0:
                i1 = line.find("'")
                path = line[i1 + 1: line.find("'", i1 + 1)]
            elif len(fields) > 2:
                tool = os.path.basename(fields[1])
                if tool == "sox":
                    path = fields[2]
                elif tool == "flac":
                    path = fields[-1]
                else:
                    raise RuntimeError(f"Unknown wav.scp format with {fields[1]}")
            else:
                path = fields[1]
            if "$" in path:
                path = os.path.expandvars(path)
            wav[wavid] = path
    return wav


@dataclass
class KaldiFolder:
    id2text: Dict[str, str]
    id2spk: Dict[str, str]
    id2dur: Dict[str, float]
    id2seg: Dict[str, Tuple[str, float, float]]   # id -> (wav id, start, end)
    wav2path: Dict[str, str]
    has_segments: bool


def read_kaldi_folder(dirin: str, glue_starting_punctuation_to_previous: bool = True) -> KaldiFolder:
    """tools/align_audio_transcript.py:190-238."""
    id2text: Dict[str, str] = {}
    previous_id = None
    with open(os.path.join(dirin, "text")) as f:
        for line in f:
            id_text = line.strip().split(" ", 1)
            if len(id_text) == 1:
                continue
            uid, text = id_text
            text = text.strip()
            if (glue_starting_punctuation_to_previous and previous_id and text and text[0] in ".,:;?!"
                    and (len(text) == 1 or text[1] in " ") and id2text[previous_id][-1] not in ".,:;?!"):
                id2text[previous_id] += text[0]
                text = text[1:].strip()
            if not text:
                continue
            id2text[uid] = text
            previous_id = uid
    with open(os.path.join(dirin, "utt2spk")) as f:
        id2spk = dict(line.strip().split() for line in f if line.strip())
    id2dur = {}
    with open(os.path.join(dirin, "utt2dur")) as f:
        for line in f:
            if line.strip():
                uid, dur = line.strip().split(" ")
                id2dur[uid] = float(dur)
    seg_path = os.path.join(dirin, "segments")
    has_segments = os.path.isfile(seg_path)
    if has_segments:
        id2seg = {}
        with open(seg_path) as f:
            for line in f:
                if line.strip():
                    uid, wav_, start_, end_ = line.strip().split(" ")
                    id2seg[uid] = (wav_, float(start_), float(end_))
    else:
        id2seg = {uid: (uid, 0, id2dur[uid]) for uid in id2dur}
    return KaldiFolder(id2text, id2spk, id2dur, id2seg, parse_kaldi_wavscp(os.path.join(dirin, "wav.scp")),
                       has_segments)


class KaldiCutWriter:
    """Appends to text / utt2spk / utt2dur / segments of `dirout` with the reference's line formats and flushes
    after every cut, so that an interrupted run can be resumed (:242-252, :419-423)."""

    def __init__(self, dirout: str):
        os.makedirs(dirout, exist_ok=True)
        self._f = {k: open(os.path.join(dirout, k), "a") for k in ("text", "utt2spk", "utt2dur", "segments")}

    def write_original(self, uid: str, transcript: str, spk: str, dur, seg: Tuple[str, float, float]) -> None:
        """:301-307 -- an utterance that is short enough already."""
        self._f["text"].write(f"{uid} {transcript}\n")
        self._f["utt2spk"].write(f"{uid} {spk}\n")
        self._f["utt2dur"].write(f"{uid} {dur}\n")
        self._f["segments"].write(f"{uid} {seg[0]} {seg[1]} {seg[2]}\n")

    def write_cut(self, new_id: str, transcript: str, spk: str, wavid: str, new_start: float, new_end: float) -> None:
        """:419-422."""
        self._f["text"].write(f"{new_id} {transcript}\n")
        self._f["utt2spk"].write(f"{new_id} {spk}\n")
        self._f["utt2dur"].write(f"{new_id} {new_end - new_start:.3f}\n")
        self._f["segments"].write(f"{new_id} {wavid} {new_start:.3f} {new_end:.3f}\n")
        for f in self._f.values():
            f.flush()

    def close(self) -> None:
        for f in self._f.values():
            f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ------------------------------------------------------------------------------------------- word packing
def regroup_isolated_punctuation(words: Sequence[str]) -> List[str]:
    """:309-315 -- a "word" made of punctuation only is glued (with a space) to the previous word."""
    out: List[str] = []
    for w in words:
        if out and re.sub(rf"[ {re.escape(PUNCTUATION)}]", "", w) == "":
            out[-1] += " " + w
        else:
            out.append(w)
    return out


@dataclass
class Cut:
    index: int          # the NN of `<id>_cutNN` (counts skipped cuts as well, like the reference)
    start: float        # seconds, relative to the wav file
    end: float
    transcript: str
    written: bool       # False: null / negative duration, or too long with skip_too_long


def pack_words(word_spans: Sequence[Tuple[int, int]], words: Sequence[str], num_frames: int, ratio: float,
               start: float, max_duration: float, refine_timestamps: float = 0, skip_too_long: bool = False) -> List[Cut]:
    """:392-435.  word_spans: (start_frame, end_frame) of every word (Segment.start / .end of word_segments),
    `ratio` seconds per frame (:371), `start` the offset of the utterance in its wav file."""
    assert len(word_spans) == len(words), f"{len(word_spans)} != {len(words)}"
    spans = [[int(s), int(e)] for s, e in word_spans]
    if spans and not refine_timestamps:      # :399-401
        spans[0][0] = 0
        spans[-1][1] = num_frames
    cuts: List[Cut] = []
    state = {"index": 1, "first": 0.0, "last": 0.0, "text": ""}

    def add_segment():                       # :383-425
        new_start, new_end = start + state["first"], start + state["last"]
        ok = not (state["last"] <= state["first"]) and not (new_end - new_start > max_duration and skip_too_long)
        cuts.append(Cut(state["index"], new_start, new_end, state["text"], ok))
        state["index"] += 1
        state["first"] = state["last"]
        state["text"] = ""

    for i, (span, word) in enumerate(zip(spans, words)):
        if word.strip() in PUNCTUATION:      # :396-397, :403-404 (substring test on the punctuation string)
            span[1] = span[0]
        if refine_timestamps and i == 0:     # :407-408
            state["first"] = state["last"] = span[0] * ratio
        end = span[1] * ratio
        if end - state["first"] > max_duration and state["text"]:
            add_segment()
        state["last"] = end
        if state["text"]:
            state["text"] += " "
        state["text"] += word
    if state["text"]:
        state["last"] = spans[-1][1] * ratio
        add_segment()
    return cuts


# ------------------------------------------------------------------------------------------- long audio
def chunked_emission(infer_sub: Callable[[int, int], "object"], n_samples: int, max_len: int):
    """Emissions of an audio longer than the acoustic model can take at once (SURVEY 8 f-4:
    ssak/infer/transformers_infer.py:259-265, torchaudio_infer.py:49-56): `infer_sub(i, j)` returns the [1, T_ij, V]
    (or [T_ij, V]) output for the samples [i, j); consecutive windows of `max_len` samples, concatenated along the
    frame axis -- the seams are where the 10-minute alignments of config C3 come from."""
    import torch
    if n_samples <= max_len:
        return infer_sub(0, n_samples)
    parts = [infer_sub(i, min(i + max_len, n_samples)) for i in range(0, n_samples, max_len)]
    return torch.cat(parts, dim=1 if parts[0].dim() == 3 else 0)


# ------------------------------------------------------------------------------------------- driver
def cut_kaldi_folder(dirin: str, dirout: str, emission_fn: Callable, labels: Sequence[str], blank_id: int,
                     max_duration: float = 30.0, min_duration: float = 0.005, refine_timestamps: float = 0,
                     batch_size: int = 16, normalize: Optional[Callable[[str], str]] = None,
                     skip_too_long: bool = False) -> Dict[str, int]:
    """Cut the long utterances of a Kaldi folder at word boundaries (tools/align_audio_transcript.py:257-436),
    `batch_size` utterances per aligner launch.

    emission_fn(utt_id, wav_path, start, end) -> (emission [T,V] CUDA tensor of log-probabilities,
    audio_seconds): the reference's load_audio + compute_logprobas (:322, ssak/utils/align_transcriptions.py:304).
    Returns counters {kept, cut, removed}."""
    from .align import compute_alignments
    folder = read_kaldi_folder(dirin)
    stats = {"kept": 0, "cut": 0, "removed": 0}
    pending: List[tuple] = []

    def flush(writer: KaldiCutWriter):
        if not pending:
            return
        results = compute_alignments([p[0] for p in pending], [p[1] for p in pending], labels, blank_id,
                                     first_as_garbage=bool(refine_timestamps))
        for (em, words, uid, wavid, start, seconds), res in zip(pending, results):
            if res is None:                  # "Failed to align": the reference skips the utterance (:340-345)
                stats["removed"] += 1
                continue
            _, word_segments = res
            num_frames = int(em.shape[0])
            ratio = seconds / num_frames     # = len(audio) / (num_frames * sample_rate), :371
            for cut in pack_words([(w.start, w.end) for w in word_segments], words, num_frames, ratio, start,
                                  max_duration, refine_timestamps, skip_too_long):
                if cut.written:
                    writer.write_cut(f"{uid}_cut{cut.index:02}", cut.transcript, folder.id2spk[uid], wavid,
                                     cut.start, cut.end)
                    stats["cut"] += 1
        pending.clear()

    with KaldiCutWriter(dirout) as writer:
        for uid, dur in folder.id2dur.items():
            if uid not in folder.id2text:
                continue
            transcript = folder.id2text[uid] if normalize is None else normalize(folder.id2text[uid])
            if not transcript or dur <= min_duration:
                stats["removed"] += 1
                continue
            wavid, start, end = folder.id2seg[uid]
            if dur <= max_duration and not refine_timestamps:
                flush(writer)                # keep the output in input order
                writer.write_original(uid, transcript, folder.id2spk[uid], folder.id2dur[uid], folder.id2seg[uid])
                stats["kept"] += 1
                continue
            words = regroup_isolated_punctuation(transcript.split())
            if refine_timestamps:
                start = max(0, start - refine_timestamps)
                end = end + refine_timestamps
            em, seconds = emission_fn(uid, folder.wav2path[wavid], start, end)
            pending.append((em, words, uid, wavid, start, seconds))
            if len(pending) >= batch_size:
                flush(writer)
        flush(writer)
    return stats
