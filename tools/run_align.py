import os, sys, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200
from ssak_b200.synth import align_batch
cfgs = {"m1": (148, 3000, 50, 100, 120, 3000), "m2": (148, 3000, 50, 220, 250, 3000), "m4": (148, 3000, 50, 480, 500, 3000), "m8": (148, 3000, 50, 1000, 1020, 3000), "c2shape": (64, 1500, 50, 200, 400, 1200), "c5": (512, 750, 1024, 100, 200, 600), "c3lite": (4, 30000, 50, 7600, 8000, 30000), "c3": (16, 30000, 50, 7600, 8000, 30000)}
name = sys.argv[1] if len(sys.argv) > 1 else "c2shape"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B, T, V, Lmin, Lmax, Tmin = cfgs[name]
em, toks, el, tl = align_batch(B, T, V, Lmin, Lmax, 5, Tmin=Tmin)
em_d, toks_d, el_d, tl_d = em.cuda(), toks.cuda(), el.cuda(), tl.cuda()
ts = []
for i in range(iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = ssak_b200.forced_align(em_d, toks_d, el_d, tl_d); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
cells = int((el.long() * (tl.long() + 1)).sum())
print(json.dumps({"cfg": name, "ms": statistics.mean(ts[1:]), "cells_per_s": cells / (statistics.mean(ts[1:]) * 1e-3), "ok": int((r.status == 0).sum())}))
