"""Single GPU: device time per step of ctc_loss vs sharded_ctc_loss (no process group: the collective is
skipped), and the CPU wall time per step -- shows whether the sharded wrapper is launch-bound."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200
from ssak_b200.shard import sharded_ctc_loss
from ssak_b200.synth import ctc_batch
lp, tg, il, tl = ctc_batch(64, 1500, 50, 200, 400, 1236, Tmin=1200)
lp_d, tg_d, il_d, tl_d = lp.cuda(), tg.cuda().int(), il.cuda().int(), tl.cuda().int()
def plain():
    x = lp_d.detach().requires_grad_(True)
    ssak_b200.ctc_loss(x, tg_d, il_d, tl_d, 0, "mean", True).backward()
def sharded():
    x = lp_d.detach().requires_grad_(True)
    sharded_ctc_loss(x, tg_d, il_d, tl_d, 0, "mean", True, global_batch=128).backward()
for name, fn in (("plain", plain), ("sharded", sharded)):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(50): fn()
    t_cpu = time.perf_counter() - t0
    b.record(); torch.cuda.synchronize()
    print(name, "device ms/step", a.elapsed_time(b) / 50, "cpu issue ms/step", t_cpu / 50 * 1e3)
