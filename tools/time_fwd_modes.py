"""C2: forward launch (rows saved) + backward launch, wavefront vs barrier forward (env SSAK_CTC_FWD_WAVE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200, bench
lib = ssak_b200.lib(); dev = torch.device("cuda", 0)
flush = torch.zeros(96 * 1024 * 1024, dtype=torch.float32, device=dev)
B, T, V, Lmin, Lmax, Tmin = bench.WORKLOADS["c2"]
lp, tg, il, tl, cells = bench.make_batch("c2", 99)
lp_d = lp.to(dev); off = torch.arange(B, device=dev, dtype=torch.int64) * tg.shape[1]
args = (tg.to(torch.int32).to(dev), off, il.to(torch.int32).to(dev), tl.to(torch.int32).to(dev), int(tl.max()))
for mode in sys.argv[1:] or ["0", "1"]:
    os.environ["SSAK_CTC_FWD_WAVE"] = mode
    tf, tb = bench.time_kernels(lib, dev, lp_d, *args, 5, flush)
    print("FWD_WAVE", mode, "fwd_ms", round(tf * 1e3, 4), "bwd_ms", round(tb * 1e3, 4), flush=True)
