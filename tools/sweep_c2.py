import os, sys, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200, bench
lib = ssak_b200.lib(); dev = torch.device("cuda", 0)
flush = torch.zeros(96 * 1024 * 1024, dtype=torch.float32, device=dev)
B, T, V, Lmin, Lmax, Tmin = bench.WORKLOADS["c2"]
lp, tg, il, tl, cells = bench.make_batch("c2", 99)
lp_d = lp.to(dev); off = torch.arange(B, device=dev, dtype=torch.int64) * tg.shape[1]
args = (tg.to(torch.int32).to(dev), off, il.to(torch.int32).to(dev), tl.to(torch.int32).to(dev), int(tl.max()))
for K, G, oc, ost in itertools.product((2,), (2, 4, 8), (1, 2, 4, 8), (3, 5, 8)):
    env = {"SSAK_CTC_K": K, "SSAK_CTC_G": G, "SSAK_CTC_OR_CHUNK": oc, "SSAK_CTC_OR_STAGES": ost}
    for k, v in env.items(): os.environ[k] = str(v)
    try:
        tf, tb = bench.time_kernels(lib, dev, lp_d, *args, 3, flush)
        print(json.dumps({**{k[9:]: v for k, v in env.items()}, "fwd_ms": round(tf * 1e3, 4), "bwd_ms": round(tb * 1e3, 4)}), flush=True)
    except AssertionError:
        pass
