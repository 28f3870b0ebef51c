import os, sys, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200, bench
lib = ssak_b200.lib(); dev = torch.device("cuda", 0)
flush = torch.zeros(96 * 1024 * 1024, dtype=torch.float32, device=dev)
for name in sys.argv[1:] or ["1k", "c5"]:
    B, T, V, Lmin, Lmax, Tmin = bench.WORKLOADS[name]
    lp, tg, il, tl, cells = bench.make_batch(name, 99)
    lp_d = lp.to(dev); off = torch.arange(B, device=dev, dtype=torch.int64) * tg.shape[1]
    args = (tg.to(torch.int32).to(dev), off, il.to(torch.int32).to(dev), tl.to(torch.int32).to(dev), int(tl.max()))
    for K, G, oc, ost, few in itertools.product((2, 4), (1, 2, 4), (1, 2, 4), (3,), (0, 1)):
        env = {"SSAK_CTC_K": K, "SSAK_CTC_G": G, "SSAK_CTC_OR_CHUNK": oc, "SSAK_CTC_OR_STAGES": ost, "SSAK_CTC_FEW": few}
        for k, v in env.items(): os.environ[k] = str(v)
        try:
            tf, tb = bench.time_kernels(lib, dev, lp_d, *args, 3, flush)
            print(json.dumps({"w": name, **{k[9:]: v for k, v in env.items()}, "fwd_ms": round(tf * 1e3, 3), "bwd_ms": round(tb * 1e3, 3)}), flush=True)
        except AssertionError:
            print(json.dumps({"w": name, **{k[9:]: v for k, v in env.items()}, "error": 1}), flush=True)
