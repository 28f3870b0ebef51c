"""Per-utterance detail for a list of fuzz cases, through the C ABI (forward + backward, grad_out = 1):
python tools/fuzz_diag.py "V Lmax T B planted logits seed" ...   -- path flags, likelihoods, where the gradient is off,
and the same through the log-domain kernels alone (SSAK_CTC_LIN32=0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import ssak_b200
from ssak_b200.synth import ctc_batch


def run(x0, tg, il, tl, logits, mode):
    os.environ["SSAK_CTC_LIN32"] = mode
    L = ssak_b200.lib()
    T, B, V = x0.shape
    dev = torch.device("cuda", 0)
    x = x0.to(dev).contiguous()
    tg32 = tg.to(dev, torch.int32).contiguous()
    off = torch.arange(B, device=dev, dtype=torch.int64) * tg32.shape[1]
    il32, tl32 = il.to(dev, torch.int32), tl.to(dev, torch.int32)
    lmax = int(tl.max())
    wsb = L.ssak_ctc_loss_workspace_bytes_v(T, B, V, lmax, 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    nll = torch.empty(B, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    fwd = L.ssak_ctc_logits_forward if logits else L.ssak_ctc_loss_forward
    bwd = L.ssak_ctc_logits_backward if logits else L.ssak_ctc_loss_backward
    rc = fwd(x.data_ptr(), T, B, V, x.stride(0), x.stride(1), tg32.data_ptr(), off.data_ptr(), il32.data_ptr(),
             tl32.data_ptr(), lmax, 0, 1, nll.data_ptr(), ws.data_ptr(), wsb, s)
    assert rc == 0, rc
    nll_fwd = nll.clone()
    fl0 = torch.zeros(B, dtype=torch.int32, device=dev)
    L.ssak_ctc_loss_path_flags(ws.data_ptr(), T, B, V, lmax, 1, fl0.data_ptr(), s)
    grad = torch.full_like(x, 7.0)
    go = torch.ones(B, device=dev)
    rc = bwd(go.data_ptr(), x.data_ptr(), T, B, V, x.stride(0), x.stride(1), tg32.data_ptr(), off.data_ptr(),
             il32.data_ptr(), tl32.data_ptr(), lmax, 0, 1, nll.data_ptr(), grad.data_ptr(), grad.stride(0), grad.stride(1),
             ws.data_ptr(), wsb, s)
    assert rc == 0, rc
    fl = torch.zeros(B, dtype=torch.int32, device=dev)
    L.ssak_ctc_loss_path_flags(ws.data_ptr(), T, B, V, lmax, 1, fl.data_ptr(), s)
    torch.cuda.synchronize()
    return fl0.cpu(), fl.cpu(), nll_fwd.cpu(), nll.cpu(), grad.cpu()


for case in sys.argv[1:]:
    V, Lmax, T, B, planted, logits, seed = (int(v) for v in case.split())
    lp, tg, il, tl = ctc_batch(B, T, V, 0, Lmax, seed, Tmin=1, planted=bool(planted) and V > 2)
    tl = torch.minimum(tl, torch.tensor(Lmax)); il = torch.clamp(il, 1, T)
    x0 = lp * 1.7 + 0.3 if logits else lp
    y = x0.double().requires_grad_(True)
    ref = F.ctc_loss(F.log_softmax(y, -1) if logits else y, tg, il, tl, 0, "none", False); ref.sum().backward()
    gref = torch.where(torch.isfinite(ref)[None, :, None], y.grad, torch.zeros_like(y.grad))
    print(f"=== {case}")
    for mode in ("1", "0"):
        fl0, fl, n0, n1, g = run(x0, tg, il, tl, bool(logits), mode)
        print(f"  -- SSAK_CTC_LIN32={mode}")
        for b in range(B):
            Tb, Lb = int(il[b]), int(tl[b])
            rep = int((tg[b, 1:Lb] == tg[b, :Lb - 1]).sum()) if Lb > 1 else 0
            e = (g[:, b].double() - gref[:, b]).abs()
            em = e.amax(1)
            t_bad = int(em.argmax())
            nbad = int((em > 1e-4).sum())
            print(f"  b={b} Tb={Tb} L={Lb} rep={rep} slack={Tb - Lb - rep} flags fwd={int(fl0[b])} end={int(fl[b])} "
                  f"ref={float(ref[b]):.4f} nll_fwd={float(n0[b]):.4f} nll_end={float(n1[b]):.4f} "
                  f"gerr={float(e.max()):.2e} at t={t_bad} (frames>1e-4: {nbad}, first {int((em > 1e-4).nonzero()[0]) if nbad else -1}, "
                  f"last {int((em > 1e-4).nonzero()[-1]) if nbad else -1}) ours[t]max={float(g[t_bad, b].abs().max()):.3e} "
                  f"nonfinite={int((~torch.isfinite(g[:, b])).sum())}")
