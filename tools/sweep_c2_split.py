"""C2 backward with posterior warps: gradient warps / row-ring geometry / K sweep (env overrides)."""
import os, sys, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200, bench
lib = ssak_b200.lib(); dev = torch.device("cuda", 0)
flush = torch.zeros(96 * 1024 * 1024, dtype=torch.float32, device=dev)
name = "c2"
B, T, V, Lmin, Lmax, Tmin = bench.WORKLOADS[name]
lp, tg, il, tl, cells = bench.make_batch(name, 99)
lp_d = lp.to(dev); off = torch.arange(B, device=dev, dtype=torch.int64) * tg.shape[1]
args = (tg.to(torch.int32).to(dev), off, il.to(torch.int32).to(dev), tl.to(torch.int32).to(dev), int(tl.max()))
keys = ("SSAK_CTC_K", "SSAK_CTC_G", "SSAK_CTC_OR_CHUNK", "SSAK_CTC_OR_STAGES", "SSAK_CTC_SPLIT")
for K, G, oc, ost, sp in [(2, 4, 8, 3, 1), (2, 3, 8, 3, 1), (2, 5, 8, 3, 1), (2, 6, 8, 3, 1), (2, 4, 8, 3, 1), (2, 4, 8, 3, 0)]:
    for k, v in zip(keys, (K, G, oc, ost, sp)): os.environ[k] = str(v)
    try:
        tf, tb = bench.time_kernels(lib, dev, lp_d, *args, 5, flush)
        print(json.dumps({"K": K, "G": G, "oc": oc, "ost": ost, "split": sp, "fwd_ms": round(tf * 1e3, 4), "bwd_ms": round(tb * 1e3, 4)}), flush=True)
    except Exception as e:
        print(json.dumps({"K": K, "G": G, "oc": oc, "ost": ost, "split": sp, "error": repr(e)[:80]}), flush=True)
