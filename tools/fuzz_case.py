"""Replay one loss case of the fuzz with per-utterance detail:
[SSAK_CTC_LIN32=.. SSAK_CTC_FWD_WAVE=.. ...] python tools/fuzz_case.py V Lmax T B planted logits seed"""
import os, sys
os.environ.setdefault("SSAK_CTC_LIN32", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, torch.nn.functional as F
import ssak_b200
from ssak_b200.synth import ctc_batch
from test_gpu_lin32 import _path_flags
V, Lmax, T, B, planted, logits, seed = (int(x) for x in sys.argv[1:8])
lp, tg, il, tl = ctc_batch(B, T, V, 0, Lmax, seed, Tmin=1, planted=bool(planted) and V > 2)
tl = torch.minimum(tl, torch.tensor(Lmax)); il = torch.clamp(il, 1, T)
x0 = lp * 1.7 + 0.3 if logits else lp
x = x0.cuda().requires_grad_(True)
fn = ssak_b200.ctc_loss_from_logits if logits else ssak_b200.ctc_loss
loss = fn(x, tg, il, tl, 0, "none", True); loss.sum().backward()
y = x0.double().requires_grad_(True)
ref = F.ctc_loss(F.log_softmax(y, -1) if logits else y, tg, il, tl, 0, "none", True); ref.sum().backward()
rep = [(int(((tg[b, 1:int(tl[b])] == tg[b, :int(tl[b]) - 1]).sum())) if int(tl[b]) > 1 else 0) for b in range(B)]
print("il", il.tolist()); print("tl", tl.tolist()); print("repeats", rep)
print("ours", [round(v, 4) for v in loss.detach().cpu().tolist()]); print("ref ", [round(v, 4) for v in ref.detach().tolist()])
print("grad err per utt", [f"{e:.1e}" for e in (x.grad.cpu().double() - y.grad).abs().amax(dim=(0, 2)).tolist()])
if not logits:
    fl, _, _ = _path_flags(lp, tg, il, tl)
    print("flags", fl.tolist())
