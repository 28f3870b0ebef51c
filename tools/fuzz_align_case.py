"""Replay one aligner case of the fuzz on every aligner kernel, with the column 0 of first_as_garbage computed on the
device (the default) and handed in from the CPU (torch CPU ops, what the reference runs):
python tools/fuzz_align_case.py V Lmax T B kind first_as_garbage seed"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ssak_b200
from oracle import oracle as O
from ssak_b200.synth import align_batch

V, Lmax, T, B = (int(v) for v in sys.argv[1:5])
kind, fag, seed = sys.argv[5], sys.argv[6] in ("1", "True"), int(sys.argv[7])
em, toks, el, tl = align_batch(B, T, V, 1, Lmax, seed, Tmin=1, kind=kind)
ref = [O.align(em[b, :int(el[b])].numpy(), toks[b, :int(tl[b])].tolist(), 0, fag) for b in range(B)]
c0_cpu = torch.zeros(B, T)
for b in range(B):
    if int(tl[b]) > 0:
        c0_cpu[b, :int(el[b])] = torch.from_numpy(O.garbage_col0(em[b, :int(el[b])].numpy(), int(toks[b, 0])))
c0_dev = (1 - em.cuda().gather(2, toks[:, :1].long().cuda().view(B, 1, 1).expand(B, T, 1)).squeeze(2).exp()).log().cpu()
for b in range(B):
    d = (c0_dev[b, :int(el[b])] != c0_cpu[b, :int(el[b])])
    print(f"b={b} T={int(el[b])} L={int(tl[b])}: column-0 elements that differ between CUDA and CPU transcendentals: {int(d.sum())}")
for name, knobs in (("lane", {"SSAK_ALIGN_LANE": "1"}), ("wave", {"SSAK_ALIGN_LANE": "0"}),
                    ("barrier", {"SSAK_ALIGN_LANE": "0", "SSAK_ALIGN_WAVE": "0"})):
    for k in ("SSAK_ALIGN_LANE", "SSAK_ALIGN_WAVE"):
        os.environ.pop(k, None)
    os.environ.update(knobs)
    for label, c0 in (("device col0", None), ("cpu col0", c0_cpu)):
        if not fag and c0 is not None:
            continue
        res = ssak_b200.forced_align(em.cuda(), toks, el, tl, first_as_garbage=fag, col0=c0)
        st, en, ts, status = res.starts.cpu(), res.ends.cpu(), res.t_start.cpu(), res.status.cpu()
        bad = []
        for b in range(B):
            rc, ss, se, sc, t0 = ref[b]
            Lb = int(tl[b])
            ok = (rc == 0) == (int(status[b]) == 0)
            if ok and rc == 0:
                ok = int(ts[b]) == t0 and st[b, :Lb].tolist() == ss.tolist() and en[b, :Lb].tolist() == se.tolist()
            if not ok:
                nd = int((st[b, :Lb] != torch.from_numpy(ss)).sum()) if rc == 0 and int(status[b]) == 0 else -1
                bad.append((b, rc, int(status[b]), nd))
        print(f"{name:8s} {label:12s}: mismatching utterances (b, oracle rc, status, tokens whose start differs): {bad}")
