"""Which utterances of a batch the throughput loss kernels hand back, and the gradient error either way:
python tools/lin32_diag.py [planted|random] [B] [T] [Lmin] [Lmax]   (SSAK_CTC_LIN32=1 is set here)"""
import os, sys
os.environ.setdefault("SSAK_CTC_LIN32", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from ssak_b200.synth import ctc_batch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_lin32 import _path_flags
kind = sys.argv[1] if len(sys.argv) > 1 else "random"
B, T, Lmin, Lmax = (int(x) for x in (sys.argv[2:6] + ["16", "1500", "200", "400"][len(sys.argv) - 2:]))
lp, tg, il, tl = ctc_batch(B, T, 50, Lmin, Lmax, 1236, Tmin=int(0.8 * T), planted=kind == "planted")
fl, nll, grad = _path_flags(lp, tg, il, tl)
y = lp.double().requires_grad_(True)
ref = F.ctc_loss(y, tg, il, tl, 0, "none", True)
ref.sum().backward()
err = (grad.double() - y.grad).abs().amax(dim=(0, 2))
print("flags", fl.tolist())
print("grad err", [f"{e:.1e}" for e in err.tolist()])
print("loss rel", [f"{e:.1e}" for e in ((nll.double() - ref.detach()).abs() / ref.detach().abs()).tolist()])
b = int(err.argmax())
e = (grad[:, b].double() - y.grad[:, b]).abs()
t = int(e.amax(dim=1).argmax())
print("worst utterance", b, "T", int(il[b]), "L", int(tl[b]), "frame", t, "col", int(e[t].argmax()))
print("frames with err > 1e-5:", (e.amax(dim=1) > 1e-5).nonzero().flatten().tolist()[:40], "count", int((e.amax(dim=1) > 1e-5).sum()))
print("ours", grad[t, b, :8].tolist()); print("ref ", y.grad[t, b, :8].tolist())
print("row sums ours/ref", float(grad[t, b].sum()), float(y.grad[t, b].sum()))
