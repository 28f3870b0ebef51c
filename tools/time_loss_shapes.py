"""fwd / bwd launch times of the loss on the named shapes (C2, 1k, C5), default configuration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200, bench
lib = ssak_b200.lib(); dev = torch.device("cuda", 0)
flush = torch.zeros(96 * 1024 * 1024, dtype=torch.float32, device=dev)
bench.WORKLOADS["b128"] = (128, 1500, 50, 200, 400, 1200)   # between the regimes: 256 CTAs
bench.WORKLOADS["b160"] = (160, 1500, 50, 200, 400, 1200)
for name in sys.argv[1:] or ["c2", "1k", "c5"]:
    B, T, V, Lmin, Lmax, Tmin = bench.WORKLOADS[name]
    lp, tg, il, tl, cells = bench.make_batch(name, 99)
    lp_d = lp.to(dev); off = torch.arange(B, device=dev, dtype=torch.int64) * tg.shape[1]
    tf, tb = bench.time_kernels(lib, dev, lp_d, tg.to(torch.int32).to(dev), off, il.to(torch.int32).to(dev),
                                tl.to(torch.int32).to(dev), int(tl.max()), 5, flush)
    print(name, "fwd_ms", round(tf * 1e3, 4), "bwd_ms", round(tb * 1e3, 4), flush=True)
    del lp_d
