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


def decode_chunked(infer_sub: Callable[[int, int], "object"], n_samples: int, max_len: int, tokenizer=None,
                   blank_id: int = 0, lm_decoder=None, beam_width: int = 100):
    """The decoding end of f-4 (ssak/infer/transformers_infer.py:84-85, 116-133, 259-265): window the audio, then either
    collapse the frame-wise argmax on the GPU (`ssak_ctc_greedy`) or hand the concatenated log-probabilities to the
    language-model beam search exactly as the reference does -- `lm_decoder.decode_beams(logits, beam_width)` with
    CPU numpy logits, best beam's text (pyctcdecode stays a host-side dependency of the reference; it is only
    called when a decoder is passed).  Returns (ids or text, emission)."""
    import torch
    from .greedy import ctc_greedy_decode
    emission = chunked_emission(infer_sub, n_samples, max_len)
    em2 = emission[0] if emission.dim() == 3 else emission
    if lm_decoder is not None:                                   # transformers_infer.py:116-133
        beams = lm_decoder.decode_beams(em2.detach().float().cpu().numpy(), beam_width=beam_width)
        return beams[0][0], emission
    ids = ctc_greedy_decode(em2.unsqueeze(0), torch.ones(1), blank_id=blank_id)[0]
    return (tokenizer.decode(ids) if tokenizer is not None else ids), emission


# ------------------------------------------------------------------------------------------- resume
def get_last_line(filename: str) -> str:
    """Last line of a file, '' for an empty file (tools/align_audio_transcript.py:466-471, without reading the whole
    file: the Kaldi files of a large corpus are hundreds of MB)."""
    size = os.path.getsize(filename)
    window = 1 << 16
    with open(filename, "rb") as f:
        while True:
            f.seek(max(0, size - window))
            lines = [l for l in f.read().split(b"\n") if l.strip()]
            if len(lines) > 1 or window >= size:
                return lines[-1].decode("utf8", "replace") if lines else ""
            window *= 16


def resume_point(dirout: str) -> Optional[str]:
    """tools/align_audio_transcript.py:160-177: None for a fresh output folder; otherwise the id of the last input
    utterance that was written (the `_cutNN` suffix stripped), after checking that all four files exist and end
    with the same id.  Raises the reference's RuntimeError / AssertionError."""
    if not os.path.isdir(dirout):
        return None
    for filename in ("utt2dur", "text", "utt2spk", "segments"):
        if not os.path.isfile(os.path.join(dirout, filename)):
            raise RuntimeError(f"Folder {dirout} already exists but does not contain file {filename}. Aborting "
                               "(remove the folder to retry)")
    line = get_last_line(os.path.join(dirout, "utt2dur"))
    if not line:
        return None
    last_id_complete = last_id = line.split()[0]
    if re.match(r".+_cut\d+$", last_id_complete):
        last_id = "_cut".join(last_id_complete.split("_cut")[:-1])
    for filename in ("text", "utt2spk", "segments"):
        line = get_last_line(os.path.join(dirout, filename))
        assert line and line.split()[0] == last_id_complete, (
            f"Last id {last_id_complete} in utt2dur does not match last id {line.split()[0] if line else None} in {filename}")
    return last_id


def reject_on_score(char_scores: Sequence[float], word_scores: Sequence[float], is_first_segment: bool,
                    is_last_segment: bool, is_weird: bool, can_reject_only_first_and_last: bool = True,
                    threshold: float = 0.4) -> bool:
    """:347-365 -- an utterance whose mean character score AND mean word score are below 0.4 is dropped; with
    can_reject_only_first_and_last only the first / last utterance of an audio file (and the ones with a "weird"
    duration) can be."""
    char_score = sum(char_scores) / max(len(char_scores), 1)
    word_score = sum(word_scores) / max(len(word_scores), 1)
    if max(char_score, word_score) >= threshold:
        return False
    return (not can_reject_only_first_and_last) or is_weird or is_first_segment or is_last_segment


# ------------------------------------------------------------------------------------------- driver
def cut_kaldi_folder(dirin: str, dirout: str, emission_fn: Callable, labels: Sequence[str], blank_id: int,
                     max_duration: float = 30.0, min_duration: float = 0.005, refine_timestamps: float = 0,
                     batch_size: int = 16, normalize: Optional[Callable[[str], str]] = None,
                     skip_too_long: bool = False, word_normalize: Optional[Callable[[str], str]] = None,
                     regex_rm_full: Sequence[str] = (), special_duration_meaning_tonext: Sequence[float] = (),
                     can_reject_based_on_score: bool = False, can_reject_only_first_and_last: bool = True,
                     warn: Callable[[str], None] = lambda msg: None) -> Dict[str, int]:
    """Cut the long utterances of a Kaldi folder at word boundaries (tools/align_audio_transcript.py:122-439),
    `batch_size` utterances per aligner launch.

    emission_fn(utt_id, wav_path, start, end) -> (emission [T,V] CUDA tensor of log-probabilities,
    audio_seconds): the reference's load_audio + compute_logprobas (:322, ssak/utils/align_transcriptions.py:304);
    a RuntimeError it raises drops the utterance (:322-325), like an alignment failure does (:340-345).
    normalize / word_normalize: the reference's custom_text_normalization (:267) / custom_word_normalization (:317)
    -- text normalisation stays with the reference.  Resumes an interrupted run (:160-177, :218-226): utterances up
    to the last id written to `dirout` are skipped.  Returns counters {kept, cut, removed, resumed_after}."""
    from .align import compute_alignments
    assert dirout != dirin
    last_id = resume_point(dirout)
    folder = read_kaldi_folder(dirin)
    stats = {"kept": 0, "cut": 0, "removed": 0, "resumed_after": last_id}
    ids = list(folder.id2dur.keys())
    if last_id is not None:                                      # :218-226
        try:
            index_last = ids.index(last_id)
        except ValueError:
            raise RuntimeError(f"Last processed id {last_id} not found in {dirin}/utt2dur")
        ids = ids[index_last + 1:]
        if not ids:
            warn(f"{dirout} already exists and is complete. Aborting.")
            return stats
    path_of = lambda uid: folder.wav2path[folder.id2seg[uid][0]]
    pending: List[dict] = []

    def flush(writer: KaldiCutWriter):
        if not pending:
            return
        try:
            results = compute_alignments([p["em"] for p in pending], [p["norm_words"] for p in pending], labels, blank_id,
                                         first_as_garbage=bool(refine_timestamps))
        except KeyboardInterrupt:
            raise
        except Exception as err:                                 # a batch-level failure: retry one by one (:340-345)
            results = []
            for p in pending:
                try:
                    results.append(compute_alignments([p["em"]], [p["norm_words"]], labels, blank_id,
                                                      first_as_garbage=bool(refine_timestamps))[0])
                except KeyboardInterrupt:
                    raise
                except Exception as err1:
                    warn(f"{p['uid']} removed because of alignment error: {err1}")
                    results.append(None)
        for p, res in zip(pending, results):
            uid = p["uid"]
            if res is None:                  # "Failed to align": the reference skips the utterance (:340-345)
                warn(f"{uid} removed because of alignment error: Failed to align")
                stats["removed"] += 1
                continue
            segments, word_segments = res
            if can_reject_based_on_score and reject_on_score([s.score for s in segments], [w.score for w in word_segments],
                                                             p["is_first"], p["is_last"], p["is_weird"],
                                                             can_reject_only_first_and_last):
                warn(f"{uid} removed because of score < 0.4")
                stats["removed"] += 1
                continue
            num_frames = int(p["em"].shape[0])
            ratio = p["seconds"] / num_frames     # = len(audio) / (num_frames * sample_rate), :371
            for cut in pack_words([(w.start, w.end) for w in word_segments], p["words"], num_frames, ratio, p["start"],
                                  max_duration, refine_timestamps, skip_too_long):
                if cut.written:
                    writer.write_cut(f"{uid}_cut{cut.index:02}", cut.transcript, folder.id2spk[uid], p["wavid"],
                                     cut.start, cut.end)
                    stats["cut"] += 1
        pending.clear()

    with KaldiCutWriter(dirout) as writer:
        previous_path = None
        for i_dur, uid in enumerate(ids):
            dur = folder.id2dur[uid]
            if uid not in folder.id2text:
                continue
            wavid, start, end = folder.id2seg[uid]
            path = folder.wav2path[wavid]
            is_first = previous_path != path                     # :261-262
            previous_path = path
            transcript = folder.id2text[uid] if normalize is None else normalize(folder.id2text[uid])
            if not transcript:
                stats["removed"] += 1
                continue
            if any(re.search(r"^" + rx + r"$", transcript) for rx in regex_rm_full):      # :271-280
                warn(f"{uid} removed because of regex")
                stats["removed"] += 1
                continue
            next_uid = ids[i_dur + 1] if i_dur + 1 < len(ids) else None
            is_weird = False
            if (refine_timestamps and folder.has_segments and special_duration_meaning_tonext and next_uid is not None
                    and min(abs(dur - d) for d in special_duration_meaning_tonext) < 0.0001
                    and path_of(next_uid) == path):              # :283-296: "up to the next segment"
                dur = folder.id2seg[next_uid][1] - start
                end = start + dur
                folder.id2seg[uid] = (wavid, start, end)
                is_weird = True
            if dur <= min_duration:
                stats["removed"] += 1
                continue
            if dur <= max_duration and not refine_timestamps and not can_reject_based_on_score:   # :300-307
                flush(writer)                # keep the output in input order
                writer.write_original(uid, transcript, folder.id2spk[uid], folder.id2dur[uid], folder.id2seg[uid])
                stats["kept"] += 1
                continue
            words = regroup_isolated_punctuation(transcript.split())
            norm_words = words if word_normalize is None else [word_normalize(w) for w in words]   # :317
            if refine_timestamps:
                start = max(0, start - refine_timestamps)
                end = end + refine_timestamps
            try:
                em, seconds = emission_fn(uid, path, start, end)
            except RuntimeError as err:                          # :322-325 audio loading error
                warn(f"{uid} removed because of audio loading error: {err}")
                stats["removed"] += 1
                continue
            is_last = next_uid is None or path_of(next_uid) != path                               # :353-361
            pending.append({"em": em, "words": words, "norm_words": norm_words, "uid": uid, "wavid": wavid, "start": start,
                            "seconds": seconds, "is_first": is_first, "is_last": is_last, "is_weird": is_weird})
            if len(pending) >= batch_size:
                flush(writer)
        flush(writer)
    return stats
