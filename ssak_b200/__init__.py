"""ssak_b200 -- B200 (sm_100a) kernels for the CTC lattice path of linto-ai/ssak.

Public API (same call signatures as the reference's call sites, see each module):
    ctc_loss, sb_ctc_loss, install            -- torch.nn.functional.ctc_loss replacement
    ctc_loss_from_logits                      -- log_softmax + ctc_loss in one (no log-probs in memory)
    compute_alignment, forced_align, get_trellis, backtrack,
    merge_repeats, merge_words, Point, Segment -- ssak/utils/align_transcriptions.py
    ctc_greedy_decode, argmax_ids             -- greedy CTC collapse
    cut_kaldi_folder, pack_words, read_kaldi_folder -- Kaldi-folder cutter around the batched aligner (tools/align_audio_transcript.py)
The compute lives in libssak_b200.so (C ABI: include/ssak_b200.h); there is no CPU fallback.
"""
from ._lib import LIB_PATH, SsakB200Error, lib  # noqa: F401
from .align import (AlignResult, Point, Segment, Trellis, backtrack, compute_alignment, compute_alignment_from_emission,  # noqa: F401
                    compute_alignments, word_positions, forced_align, get_trellis, loose_get_char_index, merge_repeats, merge_words,
                    segments_from_result)
from .cutter import (Cut, KaldiCutWriter, chunked_emission, cut_kaldi_folder, decode_chunked, pack_words, parse_kaldi_wavscp,  # noqa: F401
                     read_kaldi_folder, regroup_isolated_punctuation, reject_on_score, resume_point)
from .greedy import argmax_ids, ctc_greedy_decode, greedy_ids, hf_collapse  # noqa: F401
from .loss import ctc_loss, ctc_loss_from_logits, ctc_neg_log_likelihood, install, sb_ctc_loss, uninstall  # noqa: F401

__version__ = "0.1.0"
