"""One loss step (forward + backward) of a benchmark workload, for ncu: python tools/prof_step.py c2|1k|c5|c4 [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssak_b200
from bench import make_batch
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lp, tg, il, tl, cells = make_batch(name, 99 if name != "c4" else 1238)
x0 = lp.cuda(); tgd, ild, tld = tg.cuda().int(), il.cuda().int(), tl.cuda().int()
for _ in range(steps):
    x = x0.detach().requires_grad_(True)
    ssak_b200.ctc_loss(x, tgd, ild, tld, 0, "mean", True).backward()
torch.cuda.synchronize()
print("ok", name, cells)
