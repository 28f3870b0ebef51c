"""Time the loss forward / backward launches for several (K, warps) settings (GPU box only)."""
import os, sys, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssak_b200
import bench

def main():
    lib = ssak_b200.lib()
    dev = torch.device("cuda", 0)
    flush = torch.zeros(96 * 1024 * 1024, dtype=torch.float32, device=dev)
    names = sys.argv[1:] or ["c2", "1k", "c5"]
    for name in names:
        B, T, V, Lmin, Lmax, Tmin = bench.WORKLOADS[name]
        lp, tg, il, tl, cells = bench.make_batch(name, 99)
        lp_d = lp.to(dev)
        off = torch.arange(B, device=dev, dtype=torch.int64) * tg.shape[1]
        args = (tg.to(torch.int32).to(dev), off, il.to(torch.int32).to(dev), tl.to(torch.int32).to(dev), int(tl.max()))
        for K in (1, 2, 4, 8):
            os.environ["SSAK_CTC_K"] = str(K)
            try:
                tf, tb = bench.time_kernels(lib, dev, lp_d, *args, 5, flush)
                print(json.dumps({"workload": name, "K": K, "fwd_ms": round(tf * 1e3, 4), "bwd_ms": round(tb * 1e3, 4),
                                  "cells_per_s": cells / (tf + tb)}), flush=True)
            except AssertionError as e:
                print(json.dumps({"workload": name, "K": K, "error": "unsupported"}), flush=True)
        os.environ.pop("SSAK_CTC_K", None)

if __name__ == "__main__":
    main()
