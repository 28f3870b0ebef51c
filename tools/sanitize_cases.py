"""Small invocations of every kernel family, run under compute-sanitizer (tools/run_sanitizer.sh):
wavefront aligner with S > 1 CTAs per utterance (cross-CTA seams), barrier aligner, loss forward wavefront +
backward with posterior warps (latency regime), loss in the many-CTA shape, logits entry points, greedy."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ssak_b200
from ssak_b200.synth import align_batch, ctc_batch

which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.cuda.set_device(0)
if which in ("all", "align"):
    em, toks, el, tl = align_batch(2, 420, 30, 380, 400, 1, Tmin=410)
    r = ssak_b200.forced_align(em.cuda(), toks, el, tl)                      # wave, S = 3 (cooperative launch)
    assert (r.status == 0).all()
    os.environ["SSAK_ALIGN_WAVE"] = "0"
    r = ssak_b200.forced_align(em.cuda(), toks, el, tl)                      # barrier kernel
    assert (r.status == 0).all()
    del os.environ["SSAK_ALIGN_WAVE"]
    r = ssak_b200.forced_align(em.cuda(), toks, el, tl, first_as_garbage=True)
    print("align ok")
if which in ("all", "loss"):
    lp, tg, il, tl = ctc_batch(3, 120, 30, 20, 50, 2, Tmin=90)
    x = lp.cuda().requires_grad_(True)
    ssak_b200.ctc_loss(x, tg, il, tl, 0, "mean", True).backward()           # wave forward + SPLIT backward
    os.environ["SSAK_CTC_FEW"] = "0"
    y = lp.cuda().requires_grad_(True)
    ssak_b200.ctc_loss(y, tg, il, tl, 0, "mean", True).backward()           # many-CTA shape
    del os.environ["SSAK_CTC_FEW"]
    assert (x.grad - y.grad).abs().max() < 1e-5
    z = (lp * 1.3 + 1).cuda().requires_grad_(True)
    ssak_b200.ctc_loss_from_logits(z, tg, il, tl, 0, "mean", True).backward()
    print("loss ok")
if which in ("all", "greedy"):
    em, toks, el, tl = align_batch(3, 100, 40, 5, 20, 3)
    out = ssak_b200.ctc_greedy_decode(em.cuda(), torch.ones(3), blank_id=0)
    print("greedy ok", [len(o) for o in out])
torch.cuda.synchronize()
