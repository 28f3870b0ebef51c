"""Run the reference's OWN alignment functions, unmodified, straight from /root/reference.

TEST INFRASTRUCTURE ONLY (see oracle/ssak_oracle.c).  `ssak.utils.align_transcriptions` cannot
be imported in this image (matplotlib, speechbrain, whisper ... are missing), so the module is
parsed with `ast` and only the nodes that make up the path are executed:
    USE_MAX, USE_CHAR_REPEATED (:24-25), get_trellis (:27-70), Point (:72-76),
    backtrack (:79-123), Segment (:126-138), merge_repeats (:141-157), merge_words (:159-173)
in a namespace holding just `torch` and `dataclass`.  No reference source is copied into the
repository; /root/reference only exists in the build container, never on the GPU box, so
callers must guard with `available()`.  Used by tests/golden/make_golden.py (fixture
generation) and by tests/test_oracle.py (live pin of the C restatement).
"""
from __future__ import annotations

import ast
import os

REFERENCE_ROOT = os.environ.get("SSAK_REFERENCE_ROOT", "/root/reference")
_ALIGN_PY = os.path.join(REFERENCE_ROOT, "ssak", "utils", "align_transcriptions.py")
_WANTED_DEFS = {"get_trellis", "backtrack", "merge_repeats", "merge_words", "Point", "Segment",
                "loose_get_char_index"}
_WANTED_ASSIGNS = {"USE_MAX", "USE_CHAR_REPEATED", "MISSING_LABELS"}
_ns = None


def available() -> bool:
    return os.path.isfile(_ALIGN_PY)


def namespace() -> dict:
    """Namespace with the reference's get_trellis/backtrack/merge_repeats/merge_words/Point/Segment."""
    global _ns
    if _ns is None:
        import torch
        from dataclasses import dataclass
        with open(_ALIGN_PY, "r", encoding="utf-8") as f:
            tree = ast.parse(f.read(), filename=_ALIGN_PY)
        keep = []
        for node in tree.body:
            if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in _WANTED_DEFS:
                keep.append(node)
            elif isinstance(node, ast.Assign) and any(
                    isinstance(t, ast.Name) and t.id in _WANTED_ASSIGNS for t in node.targets):
                keep.append(node)
        mod = ast.Module(body=keep, type_ignores=[])
        import unicodedata

        def transliterate(c):  # ssak/utils/text_basic.py:191-196 (one line, needed by loose_get_char_index)
            return unicodedata.normalize("NFKD", c).encode("ascii", "ignore").decode("ascii")

        ns = {"torch": torch, "dataclass": dataclass, "transliterate": transliterate,
              "hashmd5": lambda d: str(sorted(d.items()))}
        exec(compile(mod, _ALIGN_PY, "exec"), ns)
        missing = (_WANTED_DEFS | _WANTED_ASSIGNS) - set(ns)
        if missing:
            raise RuntimeError(f"reference extraction incomplete: {sorted(missing)}")
        _ns = ns
    return _ns


def align(emission, tokens, blank_id=0, first_as_garbage=False, want_trellis=False):
    """Reference get_trellis -> backtrack -> merge_repeats on a torch CPU tensor [T,V].

    Returns dict(status, path=[(token_index, time_index, score)], segments=[(token, start, end,
    score)], t_start, trellis?) with status -1 when the reference raises "Failed to align"."""
    import torch
    ns = namespace()
    em = torch.as_tensor(emission, dtype=torch.float32)
    toks = [int(x) for x in tokens]
    trellis = ns["get_trellis"](em, toks, blank_id=blank_id, first_as_garbage=first_as_garbage)
    out = {"status": 0, "path": [], "segments": [],
           "t_start": int(torch.argmax(trellis[:, trellis.size(1) - 1]).item())}
    if want_trellis:
        out["trellis"] = trellis.numpy().copy()
    try:
        path = ns["backtrack"](trellis, em, toks, blank_id=blank_id)
    except RuntimeError as e:
        if "Failed to align" not in str(e):
            raise
        out["status"] = -1
        return out
    out["path"] = [(p.token_index, p.time_index, p.score) for p in path]
    segs = ns["merge_repeats"](list(range(len(toks))), path)
    out["segments"] = [(s.label, s.start, s.end, s.score) for s in segs]
    return out


# --------------------------------------------------------------------------------------------------
# the word-packing step of tools/align_audio_transcript.py (:383-435), i.e. `add_segment` and the
# loop that packs aligned words into cuts of at most `max_duration` seconds.  It lives inside a 300-line function
# with heavy imports, so the block is sliced out of the source text (never copied into the repository), dedented
# and executed on stand-in file objects; _punctuation comes from ssak/utils/text_basic.py:15-16 the same way.
_TOOL_PY = os.path.join(REFERENCE_ROOT, "tools", "align_audio_transcript.py")
_TEXT_BASIC_PY = os.path.join(REFERENCE_ROOT, "ssak", "utils", "text_basic.py")
_pack_code = None
_punct = None


def cutter_available() -> bool:
    return os.path.isfile(_TOOL_PY) and os.path.isfile(_TEXT_BASIC_PY)


def reference_punctuation() -> str:
    global _punct
    if _punct is None:
        import string
        with open(_TEXT_BASIC_PY, "r", encoding="utf-8") as f:
            tree = ast.parse(f.read(), filename=_TEXT_BASIC_PY)
        keep = [n for n in tree.body if isinstance(n, ast.Assign) and any(
            isinstance(t, ast.Name) and t.id in ("_punctuation_strong", "_punctuation") for t in n.targets)]
        ns = {"string": string}
        exec(compile(ast.Module(body=keep, type_ignores=[]), _TEXT_BASIC_PY, "exec"), ns)
        _punct = ns["_punctuation"]
    return _punct


def _pack_block():
    global _pack_code
    if _pack_code is None:
        import textwrap
        with open(_TOOL_PY, "r", encoding="utf-8") as f:
            lines = f.read().split("\n")
        first = next(i for i, l in enumerate(lines) if l.strip().startswith("global index, first_word_start"))
        last = next(i for i in range(first, len(lines)) if lines[i].strip() == "idx_processed += 1")
        _pack_code = compile(textwrap.dedent("\n".join(lines[first:last])), _TOOL_PY + ":pack", "exec")
    return _pack_code


def pack_words(word_spans, words, num_frames, ratio, utt_id, wavid, spk, start, max_duration,
               refine_timestamps=0, skip_warnings=False):
    """Run the reference's packing block.  word_spans: [(start_frame, end_frame)] per word.
    -> dict(text, utt2spk, utt2dur, segments: the lines the reference appends to the Kaldi files; warnings)."""
    import io
    ns_align = namespace()
    segs = [ns_align["Segment"](w, int(s), int(e), 1.0) for w, (s, e) in zip(words, word_spans)]
    warnings = []

    class _Log:
        def warning(self, m):
            warnings.append(m)

        def info(self, m):
            pass

    files = {k: io.StringIO() for k in ("f_text", "f_utt2spk", "f_utt2dur", "f_segments")}
    ns = dict(files)
    ns.update(id=utt_id, wavid=wavid, id2spk={utt_id: spk}, start=start, end=0.0, max_duration=max_duration,
              skip_warnings=skip_warnings, verbose=False, logger=_Log(), do_flush=lambda: None, debug_folder=None,
              word_segments=segs, all_words=list(words), num_frames=num_frames, ratio=ratio,
              refine_timestamps=refine_timestamps, _punctuation=reference_punctuation())
    exec(_pack_block(), ns)
    return {"text": files["f_text"].getvalue(), "utt2spk": files["f_utt2spk"].getvalue(),
            "utt2dur": files["f_utt2dur"].getvalue(), "segments": files["f_segments"].getvalue(),
            "warnings": len(warnings)}


def resume_point(dirout):
    """Run the reference's own resume block (tools/align_audio_transcript.py:160-177, with its get_last_line) on an
    output folder -> last_id (None for a fresh or empty folder).  Raises what the reference raises."""
    import re
    import textwrap
    with open(_TOOL_PY, "r", encoding="utf-8") as f:
        src = f.read()
    lines = src.split("\n")
    first = next(i for i, l in enumerate(lines) if l.strip() == "last_id = None")
    last = next(i for i in range(first, len(lines)) if lines[i].strip().startswith("os.makedirs(dirout"))
    tree = ast.parse(src, filename=_TOOL_PY)
    gll = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_last_line"]

    class _Log:
        def warning(self, m):
            pass

    ns = {"os": os, "re": re, "logger": _Log(), "dirout": dirout}
    exec(compile(ast.Module(body=gll, type_ignores=[]), _TOOL_PY, "exec"), ns)
    exec(compile(textwrap.dedent("\n".join(lines[first:last])), _TOOL_PY + ":resume", "exec"), ns)
    return ns["last_id"]
