"""One batched forced alignment of a benchmark shape, for ncu: python tools/prof_align.py [B] [T] [Lmin] [Lmax] [V]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200
from ssak_b200.synth import align_batch
B, T, Lmin, Lmax, V = (int(x) for x in (sys.argv[1:6] + ["1024", "1500", "200", "400", "50"][len(sys.argv) - 1:]))
em, toks, el, tl = align_batch(B, T, V, Lmin, Lmax, 5, Tmin=int(0.8 * T))
em_d, toks_d, el_d, tl_d = em.cuda(), toks.cuda(), el.cuda(), tl.cuda()
for _ in range(2):
    r = ssak_b200.forced_align(em_d, toks_d, el_d, tl_d)
torch.cuda.synchronize()
print("ok", int((r.status == 0).sum()))
