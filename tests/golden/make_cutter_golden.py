"""Generate tests/golden/cutter_golden.json by running the REFERENCE's word-packing block
(tools/align_audio_transcript.py:383-435, executed unmodified through oracle/ref_extract.py) on seeded random
word alignments.  Run in the build container (needs /root/reference):  python tests/golden/make_cutter_golden.py"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_extract as R  # noqa: E402


def random_case(rng):
    n = rng.randint(1, 40)
    vocab = ["bonjour", "à", "tous", "c'est", "euh", "oui", "non", "peut-être", "aujourd'hui", "et", "le", "la"]
    punct = ["!", "?", ",", ".", ":", ";", "...", "«", "»", ",.", "-", "'"]
    words = [rng.choice(punct) if rng.random() < 0.15 else rng.choice(vocab) for _ in range(n)]
    num_frames = rng.randint(n, 4000)
    bounds = sorted(rng.randint(0, num_frames) for _ in range(n + 1))
    spans = []
    for i in range(n):
        s, e = bounds[i], bounds[i + 1]
        if rng.random() < 0.1:
            e = s                                  # a word the aligner gave no frames
        spans.append((s, e))
    return dict(words=words, spans=spans, num_frames=num_frames, ratio=rng.choice([0.02, 0.0200625, 0.01]),
                start=round(rng.uniform(0, 500), 2), max_duration=rng.choice([1.0, 5.0, 15.0, 30.0]),
                refine_timestamps=rng.choice([0, 0, 0.5]), skip_warnings=rng.random() < 0.3)


def main():
    assert R.cutter_available(), "needs /root/reference"
    rng = random.Random(20240917)
    cases = []
    for _ in range(60):
        c = random_case(rng)
        c["expected"] = R.pack_words(c["spans"], c["words"], c["num_frames"], c["ratio"], "utt", "wav", "spk",
                                     c["start"], c["max_duration"], c["refine_timestamps"], c["skip_warnings"])
        cases.append(c)
    out = {"punctuation": R.reference_punctuation(), "cases": cases}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cutter_golden.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=0)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
