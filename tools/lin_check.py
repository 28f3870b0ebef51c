"""Quick check of the linear-domain loss kernels against torch CPU fp64 (and timing vs the log-domain kernels)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import ssak_b200
from ssak_b200.synth import ctc_batch

def run(lp, tg, il, tl, red="none", zi=True):
    x = lp.cuda().requires_grad_(True)
    loss = ssak_b200.ctc_loss(x, tg, il, tl, 0, red, zi)
    loss.sum().backward()
    torch.cuda.synchronize()
    return loss.detach().cpu(), x.grad.cpu()

def ref(lp, tg, il, tl, red="none", zi=True):
    y = lp.double().requires_grad_(True)
    l = F.ctc_loss(y, tg, il, tl, 0, red, zi)
    l.sum().backward()
    return l.detach(), y.grad

cases = [(3, 20, 8, 1, 5, False), (5, 50, 20, 0, 12, False), (7, 120, 50, 5, 40, True), (3, 200, 50, 60, 90, True),
         (4, 64, 1024, 3, 30, False), (2, 90, 257, 40, 44, True), (6, 700, 50, 250, 330, True), (8, 1500, 50, 200, 400, True),
         (8, 1500, 50, 200, 400, False), (4, 33, 12, 1, 9, False)]
only = os.environ.get("CASE")
for ci, (B, T, V, Lmin, Lmax, planted) in enumerate(cases):
    if only is not None and int(only) != ci:
        continue
    lp, tg, il, tl = ctc_batch(B, T, V, Lmin, Lmax, 200 + ci, Tmin=max(1, T // 2), planted=planted)
    if ci == 9:
        il[0], tl[0] = 1, 1
        il[1], tl[1] = 2, 0
        il[2], tl[2] = 3, 1
        tg[3, :3] = torch.tensor([5, 5, 5]); il[3], tl[3] = 4, 3
    rl, rg = ref(lp, tg, il, tl)
    for mode in ("1", "0"):
        os.environ["SSAK_CTC_LINEAR"] = mode
        try:
            l, g = run(lp, tg, il, tl)
        except Exception as e:
            print(ci, mode, "ERROR", repr(e)); continue
        fin = torch.isfinite(rl)
        same_fin = torch.equal(torch.isfinite(l.double()), fin)
        rel = ((l.double() - rl).abs() / rl.abs().clamp_min(1e-3))[fin].max().item() if fin.any() else 0.0
        gerr = (g.double() - rg).abs().max().item() if torch.isfinite(rg).all() else float("nan")
        print(f"case {ci} B={B} T={T} V={V} L<={Lmax} planted={planted} linear={mode}: finite-match {same_fin} loss rel {rel:.2e} grad err {gerr:.2e}", flush=True)

# timing: C2 and 1k, kernels through the torch-facing API
for name, (B, T, V, Lmin, Lmax, Tmin) in {"c2": (64, 1500, 50, 200, 400, 1200), "1k": (1024, 1500, 50, 200, 400, 1200)}.items():
    if only is not None:
        break
    lp, tg, il, tl = ctc_batch(B, T, V, Lmin, Lmax, 99, Tmin=Tmin, planted=True)
    x0 = lp.cuda(); tgd, ild, tld = tg.cuda().int(), il.cuda().int(), tl.cuda().int()
    cells = int((il * (2 * tl + 1)).sum())
    for mode in ("1", "0"):
        os.environ["SSAK_CTC_LINEAR"] = mode
        def step():
            x = x0.detach().requires_grad_(True)
            ssak_b200.ctc_loss(x, tgd, ild, tld, 0, "mean", True).backward()
        for _ in range(3): step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): step()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"{name} linear={mode}: {ms:.3f} ms/step, {cells / ms / 1e6:.1f} Gcells/s", flush=True)
