import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200
from ssak_b200.synth import align_batch
B, T, V, Lmin, Lmax, Tmin = (512, 750, 1024, 100, 200, 600)
em, toks, el, tl = align_batch(B, T, V, Lmin, Lmax, 5, Tmin=Tmin)
em_d, toks_d, el_d, tl_d = em.cuda(), toks.cuda(), el.cuda(), tl.cuda()
flush = torch.zeros(96 * 1024 * 1024, dtype=torch.float32, device="cuda")
ts = []
for i in range(12):
    if len(sys.argv) > 1: flush.add_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = ssak_b200.forced_align(em_d, toks_d, el_d, tl_d); b.record(); torch.cuda.synchronize()
    ts.append(round(a.elapsed_time(b), 3))
print(ts)
