"""BASELINE config C1: the reference's tests/data Kaldi folder through a random-init wav2vec2-base CTC head
(French character vocabulary, ~50 symbols) on the CPU, then the REFERENCE's aligner and torch's CPU ctc_loss.

Run in the build container only (needs /root/reference, transformers, scipy).  The audio itself is not copied:
only the emissions the model produced and the reference outputs are stored (tests/golden/c1_golden.npz)."""
import os, sys
import numpy as np, torch, torch.nn.functional as F
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_extract as R
from scipy.io import wavfile
import torchaudio.transforms as TT
from transformers import Wav2Vec2Config, Wav2Vec2ForCTC

DATA = os.path.join(R.REFERENCE_ROOT, "tests", "data")
LABELS = ["<pad>", "<s>", "</s>", "<unk>", "|"] + list("abcdefghijklmnopqrstuvwxyz'-") + list("àâäéèêëîïôöùûüÿç")  # ssak/utils/kaldi.py:47
BLANK = 0

def load_wav(path, start=None, end=None, sr=16000):
    rate, x = wavfile.read(path)
    x = x.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1)
    x = x / max(1.0, np.abs(x).max())
    t = torch.from_numpy(x)
    if rate != sr:
        t = TT.Resample(rate, sr)(t)
    if start is not None:
        t = t[int(start * sr): int(end * sr)]
    return t

def main():
    assert R.available()
    torch.manual_seed(1234)
    cfg = Wav2Vec2Config(vocab_size=len(LABELS), pad_token_id=BLANK, ctc_loss_reduction="mean", ctc_zero_infinity=True)
    model = Wav2Vec2ForCTC(cfg).eval()
    text = dict(l.strip().split(" ", 1) for l in open(os.path.join(DATA, "kaldi/small/text")) if " " in l.strip())
    segs = {l.split()[0]: l.split()[1:] for l in open(os.path.join(DATA, "kaldi/small/segments"))}
    wavs = {"toy_bonjour": "bonjour.wav", "toy_bonjour2": "bonjour 8k.wav", "salledebain_sch_13": "tcof2channels.wav"}
    dic = {c: i for i, c in enumerate(LABELS)}
    utts = []
    for utt, (rec, st, en) in segs.items():
        if rec not in wavs or utt not in text:
            continue
        audio = load_wav(os.path.join(DATA, "audio", wavs[rec]), float(st), float(en))
        words = [w for w in text[utt].lower().split() if not w.startswith("<")]
        chars = "|".join(words)
        toks = [dic[c] for c in chars if c in dic]
        with torch.no_grad():
            logits = model(audio.unsqueeze(0)).logits[0]
        em = F.log_softmax(logits, dim=-1, dtype=torch.float32)
        utts.append((utt, em, toks))
        print(utt, "frames", em.shape[0], "tokens", len(toks))
    flat = {"n": np.array(len(utts)), "labels": np.array(LABELS)}
    Tmax, Lmax = max(e.shape[0] for _, e, _ in utts), max(len(t) for _, _, t in utts)
    B = len(utts)
    lp = torch.zeros(Tmax, B, len(LABELS))
    tg = torch.zeros(B, Lmax, dtype=torch.long)
    il, tl = torch.zeros(B, dtype=torch.long), torch.zeros(B, dtype=torch.long)
    for i, (utt, em, toks) in enumerate(utts):
        for fag in (False, True):
            ref = R.align(em.numpy(), toks, BLANK, fag)
            k = f"u{i}_g{int(fag)}_"
            flat[k + "status"] = np.array(ref["status"]); flat[k + "t_start"] = np.array(ref["t_start"])
            flat[k + "seg_start"] = np.array([s[1] for s in ref["segments"]], np.int32)
            flat[k + "seg_end"] = np.array([s[2] for s in ref["segments"]], np.int32)
            flat[k + "seg_score"] = np.array([s[3] for s in ref["segments"]], np.float64)
        flat[f"u{i}_name"] = np.array(utt); flat[f"u{i}_emission"] = em.numpy(); flat[f"u{i}_tokens"] = np.array(toks, np.int32)
        lp[: em.shape[0], i] = em; lp[em.shape[0]:, i] = em[-1]
        tg[i, : len(toks)] = torch.tensor(toks); il[i] = em.shape[0]; tl[i] = len(toks)
    for dt, key in ((torch.float32, "f32"), (torch.float64, "f64")):
        x = lp.to(dt).detach().clone().requires_grad_(True)
        loss = F.ctc_loss(x, tg, il, tl, blank=BLANK, reduction="mean", zero_infinity=True)   # the HF configuration
        loss.backward()
        flat[f"loss_{key}"] = loss.detach().numpy(); flat[f"grad_{key}"] = x.grad.numpy()
    flat["log_probs"] = lp.numpy(); flat["targets"] = tg.numpy(); flat["input_lengths"] = il.numpy(); flat["target_lengths"] = tl.numpy()
    np.savez_compressed(os.path.join(HERE, "c1_golden.npz"), **flat)
    print("saved", B, "utterances; loss", float(flat["loss_f64"]))

if __name__ == "__main__":
    main()
