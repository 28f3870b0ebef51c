"""Forward / backward launch-group times of a workload on both loss kernel families + hand-back count:
python tools/time_loss_paths.py c5|1k|c4|c2"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200, bench
name = sys.argv[1] if len(sys.argv) > 1 else "c5"
lib = ssak_b200.lib(); dev = torch.device("cuda", 0)
if "," in name:   # B,T,V,Lmin,Lmax,Tmin
    from ssak_b200.synth import ctc_batch
    Bq, Tq, Vq, Lmin, Lmax, Tmin = (int(x) for x in name.split(","))
    lp, tg, il, tl = ctc_batch(Bq, Tq, Vq, Lmin, Lmax, 99, Tmin=Tmin, planted=False)
    cells = int((il * (2 * tl + 1)).sum())
else:
    lp, tg, il, tl, cells = bench.make_batch(name, 99 if name != "c4" else 1238)
T, B, V = lp.shape
lp_d = lp.to(dev); off = torch.arange(B, device=dev, dtype=torch.int64) * tg.shape[1]
tg32, il32, tl32 = tg.to(torch.int32).to(dev), il.to(torch.int32).to(dev), tl.to(torch.int32).to(dev)
flush = torch.zeros(96 * 1024 * 1024, dtype=torch.float32, device=dev)
for mode in ("1", "0"):
    os.environ["SSAK_CTC_LIN32"] = mode
    tf, tb, wsb = bench.time_kernels(lib, dev, lp_d, tg32, off, il32, tl32, int(tl.max()), 7, flush)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev); nll = torch.empty(B, device=dev); grad = torch.empty_like(lp_d)
    go = torch.ones(B, device=dev); s = torch.cuda.current_stream().cuda_stream; Lm = int(tl.max())
    assert lib.ssak_ctc_loss_forward(lp_d.data_ptr(), T, B, V, lp_d.stride(0), lp_d.stride(1), tg32.data_ptr(), off.data_ptr(), il32.data_ptr(), tl32.data_ptr(), Lm, 0, 1, nll.data_ptr(), ws.data_ptr(), wsb, s) == 0
    assert lib.ssak_ctc_loss_backward(go.data_ptr(), lp_d.data_ptr(), T, B, V, lp_d.stride(0), lp_d.stride(1), tg32.data_ptr(), off.data_ptr(), il32.data_ptr(), tl32.data_ptr(), Lm, 0, 1, nll.data_ptr(), grad.data_ptr(), grad.stride(0), grad.stride(1), ws.data_ptr(), wsb, s) == 0
    fl = torch.empty(B, dtype=torch.int32, device=dev)
    lib.ssak_ctc_loss_path_flags(ws.data_ptr(), T, B, V, Lm, 1, fl.data_ptr(), s)
    torch.cuda.synchronize()
    print(name, "SSAK_CTC_LIN32 =", mode, "fwd ms", round(tf * 1e3, 4), "bwd ms", round(tb * 1e3, 4), "ws MB", wsb >> 20, "handed back", int((fl != 0).sum()), "cells/s %.3e" % (cells / (tf + tb)), flush=True)
