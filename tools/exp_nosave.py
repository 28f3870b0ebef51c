import os, sys, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200, bench
lib = ssak_b200.lib(); dev = torch.device("cuda", 0)
for name in ("c2", "1k"):
    B, T, V, Lmin, Lmax, Tmin = bench.WORKLOADS[name]
    lp, tg, il, tl, cells = bench.make_batch(name, 99)
    lp_d = lp.to(dev); off = torch.arange(B, device=dev, dtype=torch.int64) * tg.shape[1]
    tg32, il32, tl32 = tg.to(torch.int32).to(dev), il.to(torch.int32).to(dev), tl.to(torch.int32).to(dev)
    Lm = int(tl.max())
    for save in (1, 0):
        ws_bytes = lib.ssak_ctc_loss_workspace_bytes(T, B, Lm, save)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev); nll = torch.empty(B, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        ts = []
        for i in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = lib.ssak_ctc_loss_forward(lp_d.data_ptr(), T, B, V, lp_d.stride(0), lp_d.stride(1), tg32.data_ptr(), off.data_ptr(), il32.data_ptr(), tl32.data_ptr(), Lm, 0, save, nll.data_ptr(), ws.data_ptr(), ws_bytes, s)
            b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(name, "save", save, "fwd_ms", round(statistics.mean(ts[2:]), 4), flush=True)
