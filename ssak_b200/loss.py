"""CTC loss with torch.nn.functional.ctc_loss's signature, on the sm_100a kernels.

Reference call sites this drops into (paths relative to the reference repository):
  * HF wav2vec2: site-packages/transformers/models/wav2vec2/modeling_wav2vec2.py:1727-1736,
    configured at ssak/train/transformers/wav2vec_train.py:313-325 (reduction="mean",
    zero_infinity=True, 1-D concatenated int64 targets, transposed [T,B,V] view)
  * SpeechBrain: ssak/train/speechbrain/wav2vec_train.py:66 through the yaml key `ctc_cost`
    (ssak/train/speechbrain/fr/hyperparameters_wav2vec_finetune_cv-fr.yaml:116-117) -> `sb_ctc_loss`
  * NeMo: ssak/train/nemo/yamls/model.yaml:3 (`ctc_reduction: mean_volume`) -> `reduction="mean_volume"`
"""
from __future__ import annotations

import torch

from . import _lib

_REDUCTIONS = ("none", "mean", "sum", "mean_volume")


def _as_length_tensor(x, B: int, device, name: str):
    """-> (int32 device tensor [B], host list or None)."""
    host = None
    if isinstance(x, torch.Tensor):
        if x.dim() == 0:
            x = x.reshape(1)
        if x.numel() != B:
            raise RuntimeError(f"{name} must be of size batch_size ({B}), got {x.numel()}")
        if x.is_floating_point():
            raise RuntimeError(f"{name} must be integral")
        if not x.is_cuda:
            host = [int(v) for v in x.tolist()]
        dev = x.to(device=device, dtype=torch.int32, non_blocking=True)
    else:
        host = [int(v) for v in x]
        if len(host) != B:
            raise RuntimeError(f"{name} must be of size batch_size ({B}), got {len(host)}")
        dev = torch.tensor(host, dtype=torch.int32, device=device)
    return dev, host


_RED_CODE = {"none": 0, "mean": 1, "sum": 2, "mean_volume": 3}
_OFFSET_CACHE = {}


def _padded_offsets(B: int, smax: int, device):
    key = (B, smax, str(device))
    t = _OFFSET_CACHE.get(key)
    if t is None:
        if len(_OFFSET_CACHE) > 64:
            _OFFSET_CACHE.clear()
        t = torch.arange(B, device=device, dtype=torch.int64) * smax
        _OFFSET_CACHE[key] = t
    return t


def _launch_backward(L, from_logits, g, log_probs, targets, tgt_off, in_len, tgt_len, max_target_len, blank,
                     zero_infinity, nll, ws, ws_bytes):
    """One backward call of the C ABI -> gradient in log_probs' (dense) layout."""
    T, B, V = log_probs.shape
    # same (dense) layout as log_probs: HF hands in a transposed [B,T,V] buffer and its
    # log_softmax backward reads the gradient in that layout
    grad = torch.empty_like(log_probs)
    if grad.stride(2) != 1:
        grad = torch.empty((T, B, V), dtype=torch.float32, device=log_probs.device)
    with torch.cuda.device(log_probs.device):
        stream = torch.cuda.current_stream().cuda_stream
        bwd = L.ssak_ctc_logits_backward if from_logits else L.ssak_ctc_loss_backward
        rc = bwd(g.data_ptr(), log_probs.data_ptr(), T, B, V, log_probs.stride(0),
                 log_probs.stride(1), targets.data_ptr(), tgt_off.data_ptr(),
                 in_len.data_ptr(), tgt_len.data_ptr(), max_target_len, blank,
                 int(zero_infinity), nll.data_ptr(), grad.data_ptr(), grad.stride(0),
                 grad.stride(1), ws.data_ptr(), ws_bytes, stream)
    _lib.check(rc, "ssak_ctc_logits_backward" if from_logits else "ssak_ctc_loss_backward")
    return grad


def _apply_upstream(ctx, grad, grad_loss, per_utterance: bool):
    """Eager mode: `grad` was computed with a unit upstream gradient; apply the real one.  First call: in place
    (rows whose factor is exactly 1 -- the usual case -- are not touched); later calls (retain_graph): out of place,
    relative to what the first call applied."""
    gl = grad_loss.detach().to(torch.float32).contiguous()
    applied = getattr(ctx, "applied", None)
    if applied is not None:
        f = gl / applied
        return grad * (f.view(1, -1, 1) if per_utterance else f)
    L = _lib.lib()
    T, B, V = grad.shape
    with torch.cuda.device(grad.device):
        rc = L.ssak_ctc_grad_scale(grad.data_ptr(), T, B, V, grad.stride(0), grad.stride(1),
                                   gl.data_ptr() if per_utterance else None, None if per_utterance else gl.data_ptr(),
                                   None, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ssak_ctc_grad_scale")
    ctx.applied = gl
    return grad


class _CTCLossFunction(torch.autograd.Function):
    """aten::_ctc_loss + the reduction of aten::ctc_loss / aten::_ctc_loss_backward on libssak_b200.so.
    forward: 3 launches (half lattices, join, reduction); backward: 1 launch (+ one scale).

    Where the throughput kernels run (ssak_ctc_loss_nll_is_provisional: batches that fill the GPU) the likelihood
    the forward call returns is provisional until the backward call has verified it, so BOTH calls run inside
    forward() -- backward with the reduction's own weights as upstream gradient -- the loss is reduced from the
    final likelihoods, and backward() only applies autograd's upstream gradient (no memory traffic when it is 1)."""

    @staticmethod
    def forward(ctx, log_probs, targets, tgt_off, in_len, tgt_len, max_target_len, blank, zero_infinity,
                reduction, from_logits=False):
        L = _lib.lib()
        T, B, V = log_probs.shape
        dev = log_probs.device
        save = bool(ctx.needs_input_grad[0])
        _require_supported(T, B, V, max_target_len)
        ws_bytes = L.ssak_ctc_loss_workspace_bytes_v(T, B, V, max_target_len, int(save))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        red = _RED_CODE[reduction]
        loss = torch.empty(B if red == 0 else (), dtype=torch.float32, device=dev)
        gscale = torch.empty(B, dtype=torch.float32, device=dev) if save else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        fwd = L.ssak_ctc_logits_forward if from_logits else L.ssak_ctc_loss_forward
        rc = fwd(log_probs.data_ptr(), T, B, V, log_probs.stride(0), log_probs.stride(1),
                 targets.data_ptr(), tgt_off.data_ptr(), in_len.data_ptr(),
                 tgt_len.data_ptr(), max_target_len, blank, int(save), nll.data_ptr(),
                 ws.data_ptr(), ws_bytes, stream)
        _lib.check(rc, "ssak_ctc_logits_forward" if from_logits else "ssak_ctc_loss_forward")
        rc = L.ssak_ctc_loss_reduce(nll.data_ptr(), tgt_len.data_ptr(), B, red, int(zero_infinity),
                                    loss.data_ptr(), _lib.ptr(gscale), stream)
        _lib.check(rc, "ssak_ctc_loss_reduce")
        ctx.eager = save and bool(L.ssak_ctc_loss_nll_is_provisional(B, V, max_target_len))
        if ctx.eager:
            grad = _launch_backward(L, from_logits, gscale, log_probs, targets, tgt_off, in_len, tgt_len,
                                    max_target_len, blank, zero_infinity, nll, ws, ws_bytes)
            rc = L.ssak_ctc_loss_reduce(nll.data_ptr(), tgt_len.data_ptr(), B, red, int(zero_infinity),
                                        loss.data_ptr(), None, stream)   # from the final likelihoods
            _lib.check(rc, "ssak_ctc_loss_reduce")
            ctx.save_for_backward(grad)
            ctx.per_utterance = red == 0
        elif save:
            ctx.save_for_backward(log_probs, targets, tgt_off, in_len, tgt_len, nll, ws, gscale)
            ctx.meta = (max_target_len, blank, zero_infinity, ws_bytes, from_logits)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        if ctx.eager:
            (grad,) = ctx.saved_tensors
            return (_apply_upstream(ctx, grad, grad_loss, ctx.per_utterance),) + (None,) * 9
        log_probs, targets, tgt_off, in_len, tgt_len, nll, ws, gscale = ctx.saved_tensors
        max_target_len, blank, zero_infinity, ws_bytes, from_logits = ctx.meta
        g = (gscale * grad_loss.to(torch.float32)).contiguous()   # [B]: upstream x d loss / d nll_b
        grad = _launch_backward(_lib.lib(), from_logits, g, log_probs, targets, tgt_off, in_len, tgt_len,
                                max_target_len, blank, zero_infinity, nll, ws, ws_bytes)
        return grad, None, None, None, None, None, None, None, None, None


def supported(T: int, B: int, V: int, max_target_len: int) -> bool:
    """True when the sm_100a kernels cover the shape (ssak_ctc_loss_supported): max target length <= 4095,
    T <= 300000, B >= 1 and a vocabulary whose rows fit the shared-memory emission ring (V <= ~2040)."""
    return bool(_lib.lib().ssak_ctc_loss_supported(int(T), int(B), int(V), int(max_target_len)))


def _require_supported(T, B, V, max_target_len):
    if not supported(T, B, V, max_target_len):
        raise _lib.SsakB200Error(
            f"ctc_loss: shape not supported by the sm_100a kernels (T={T}, B={B}, V={V}, max target length="
            f"{max_target_len}; limits: T <= 300000, B >= 1, V <= ~2040, max target length <= 4095)")


def _prepare(log_probs, targets, input_lengths, target_lengths, blank):
    """Argument checking / normalisation shared by the public entry points (torch's rules)."""
    _lib.require_cuda(log_probs, "log_probs")
    if log_probs.dim() == 2:
        log_probs = log_probs.unsqueeze(1)
        if isinstance(targets, torch.Tensor) and targets.dim() == 1:
            targets = targets.unsqueeze(0)
    if log_probs.dim() != 3:
        raise RuntimeError("log_probs must be [T, B, V] (or [T, V])")
    if log_probs.dtype != torch.float32:
        log_probs = log_probs.float()
    if log_probs.stride(2) != 1 and log_probs.size(2) > 1:
        log_probs = log_probs.contiguous()
    T, B, V = log_probs.shape
    if not (0 <= int(blank) < V):
        raise RuntimeError("blank must be in label range")
    dev = log_probs.device
    in_len, in_host = _as_length_tensor(input_lengths, B, dev, "input_lengths")
    tgt_len, tgt_host = _as_length_tensor(target_lengths, B, dev, "target_lengths")
    if not isinstance(targets, torch.Tensor):
        targets = torch.as_tensor(targets)
    if targets.is_floating_point():
        raise RuntimeError("targets must be integral")
    tg = targets
    if not tg.is_cuda and tg.numel() and tgt_host is not None:
        # host targets: labels outside the vocabulary are an argument error (device targets cannot be checked without
        # a synchronisation: the kernels then return a NaN likelihood for the utterance instead of a wrong number)
        if tg.dim() == 2:
            used = torch.arange(tg.size(1)).unsqueeze(0) < torch.as_tensor(tgt_host).unsqueeze(1)
            bad = ((tg < 0) | (tg >= V)) & used
        else:
            bad = (tg[: sum(tgt_host)] < 0) | (tg[: sum(tgt_host)] >= V)
        if bool(bad.any()):
            raise RuntimeError(f"targets must be in the label range [0, {V})")
    if tg.device != dev or tg.dtype != torch.int32 or not tg.is_contiguous():
        tg = tg.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
    if tg.dim() == 2:
        if tg.size(0) != B:
            raise RuntimeError(f"targets must have batch size {B}")
        smax = tg.size(1)
        tgt_off = _padded_offsets(B, smax, dev)
        max_target_len = smax
        if tgt_host is not None:
            if max(tgt_host, default=0) > smax:
                raise RuntimeError("Expected tensor to have size at least max(target_lengths) along dimension 1")
            max_target_len = max(tgt_host, default=0)
    elif tg.dim() == 1:
        tl64 = tgt_len.to(torch.int64)
        tgt_off = torch.cumsum(tl64, 0) - tl64
        if tgt_host is None:
            # concatenated targets with device lengths (the HF call): one sync, as torch itself does
            tgt_host = [int(v) for v in tgt_len.tolist()]
        max_target_len = max(tgt_host, default=0)
        if sum(tgt_host) > tg.numel():
            raise RuntimeError("Expected targets to hold sum(target_lengths) elements")
    else:
        raise RuntimeError("targets must be 1-D (concatenated) or 2-D (padded)")
    if in_host is not None and (max(in_host, default=0) > T or min(in_host, default=0) < 0):
        raise RuntimeError(f"Expected input_lengths to have value at most {T}, but got value "
                           f"{max(in_host)} (while checking arguments for ctc_loss)")
    if tgt_host is not None and min(tgt_host, default=0) < 0:
        raise RuntimeError("Expected target_lengths to have non-negative values")
    if tg.numel() == 0:
        tg = torch.zeros(1, dtype=torch.int32, device=dev)
    return log_probs, tg, tgt_off, in_len, tgt_len, int(max_target_len)


def ctc_neg_log_likelihood(log_probs, targets, input_lengths, target_lengths, blank=0, zero_infinity=False):
    """Per-utterance negative log-likelihood [B] (0 instead of +inf with zero_infinity), differentiable
    w.r.t. `log_probs` with torch's gradient convention.  Arguments as torch.nn.functional.ctc_loss."""
    return ctc_loss(log_probs, targets, input_lengths, target_lengths, blank, "none", zero_infinity)


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean",
             zero_infinity=False):
    """Drop-in for torch.nn.functional.ctc_loss (site-packages/torch/nn/functional.py:3042-3115).

    Extra reduction "mean_volume" = sum(nll) / sum(target_lengths) (NeMo, model.yaml:3)."""
    if reduction not in _REDUCTIONS:
        raise ValueError(f"{reduction} is not a valid value for reduction")
    unbatched = isinstance(log_probs, torch.Tensor) and log_probs.dim() == 2
    lp, tg, tgt_off, in_len, tgt_len, lmax = _prepare(log_probs, targets, input_lengths, target_lengths, blank)
    with torch.cuda.device(lp.device):
        out = _CTCLossFunction.apply(lp, tg, tgt_off, in_len, tgt_len, lmax, int(blank), bool(zero_infinity),
                                     reduction)
    return out[0] if (unbatched and reduction == "none") else out


def ctc_loss_from_logits(logits, targets, input_lengths, target_lengths, blank=0, reduction="mean",
                         zero_infinity=False):
    """`ctc_loss(log_softmax(logits, -1), ...)` without materialising the log-probabilities.

    Replaces the pair at site-packages/transformers/models/wav2vec2/modeling_wav2vec2.py:1725-1736 (and
    ssak/train/speechbrain/wav2vec_train.py:54 + :66): `logits` [T,B,V] fp32 (any strides in T and B), the
    result is differentiable w.r.t. `logits`.  One extra kernel computes the [T,B] row normalisers; the lattice
    kernels read the logits themselves, so a full write + read of the [T,B,V] log-probabilities disappears."""
    if reduction not in _REDUCTIONS:
        raise ValueError(f"{reduction} is not a valid value for reduction")
    unbatched = isinstance(logits, torch.Tensor) and logits.dim() == 2
    x, tg, tgt_off, in_len, tgt_len, lmax = _prepare(logits, targets, input_lengths, target_lengths, blank)
    with torch.cuda.device(x.device):
        out = _CTCLossFunction.apply(x, tg, tgt_off, in_len, tgt_len, lmax, int(blank), bool(zero_infinity),
                                     reduction, True)
    return out[0] if (unbatched and reduction == "none") else out


def sb_ctc_loss(log_probs, targets, input_lens, target_lens, blank_index, reduction="mean"):
    """SpeechBrain's `speechbrain.nnet.losses.ctc_loss` wrapper (the `ctc_cost` yaml hook,
    ssak/train/speechbrain/fr/hyperparameters_wav2vec_finetune_cv-fr.yaml:116-117):
    log_probs [B,T,V], relative lengths in (0,1], zero_infinity=True."""
    input_lens = (input_lens * log_probs.shape[1]).round().int()
    target_lens = (target_lens * targets.shape[1]).round().int()
    lp = log_probs.transpose(0, 1)
    if reduction == "batchmean":
        red = "sum"
    elif reduction == "batch":
        red = "none"
    else:
        red = reduction
    loss = ctc_loss(lp, targets, input_lens, target_lens, blank_index, reduction=red, zero_infinity=True)
    if reduction == "batchmean":
        return loss / targets.shape[0]
    if reduction == "batch":
        N = loss.size(0)
        return loss.view(N, -1).sum(1) / target_lens.view(N, -1).sum(1)
    return loss


_torch_ctc_loss = None
_aten_lib = None


def _covers(log_probs, targets, target_lengths) -> bool:
    """Cheap pre-check used by install(): does `ctc_loss` cover this call?  (No synchronisation: the maximum target
    length is bounded by the padded width, or by 4095 for concatenated targets with device lengths.)"""
    if log_probs.dim() not in (2, 3) or log_probs.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        return False
    T, V = log_probs.shape[0], log_probs.shape[-1]
    B = log_probs.shape[1] if log_probs.dim() == 3 else 1
    if isinstance(targets, torch.Tensor) and targets.dim() == 2:
        lmax = targets.shape[1]
        if lmax > 4095 and not (isinstance(target_lengths, torch.Tensor) and target_lengths.is_cuda):
            lmax = max((int(v) for v in target_lengths), default=0)
    elif isinstance(target_lengths, torch.Tensor) and target_lengths.is_cuda:
        lmax = 0     # unknown without a sync; _prepare finds out and ctc_loss raises beyond the limit
    else:
        lmax = max((int(v) for v in target_lengths), default=0)
    return B >= 1 and supported(T, B, V, min(lmax, 4095) if lmax <= 4095 else lmax)


def _aten_ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, zero_infinity=False):
    """aten::_ctc_loss on CUDA -> (neg_log_likelihood[B], opaque tensor handed back as `log_alpha`): the workspace
    (uint8), or -- where the forward likelihood is provisional until the backward call has run, see
    _CTCLossFunction -- the gradient for a unit upstream gradient (fp32 [T,B,V]), computed right here."""
    lp, tg, tgt_off, in_len, tgt_len, lmax = _prepare(log_probs, targets, list(input_lengths), list(target_lengths),
                                                       blank)
    L = _lib.lib()
    T, B, V = lp.shape
    _require_supported(T, B, V, lmax)
    ws_bytes = L.ssak_ctc_loss_workspace_bytes_v(T, B, V, lmax, 1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=lp.device)
    nll = torch.empty(B, dtype=torch.float32, device=lp.device)
    with torch.cuda.device(lp.device):
        rc = L.ssak_ctc_loss_forward(lp.data_ptr(), T, B, V, lp.stride(0), lp.stride(1), tg.data_ptr(),
                                     tgt_off.data_ptr(), in_len.data_ptr(), tgt_len.data_ptr(), lmax, int(blank), 1,
                                     nll.data_ptr(), ws.data_ptr(), ws_bytes,
                                     torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "ssak_ctc_loss_forward")
        if L.ssak_ctc_loss_nll_is_provisional(B, V, lmax):
            ones = torch.ones(B, dtype=torch.float32, device=lp.device)
            ws = _launch_backward(L, False, ones, lp, tg, tgt_off, in_len, tgt_len, lmax, int(blank),
                                  bool(zero_infinity), nll, ws, ws_bytes)
    return nll.to(log_probs.dtype), ws   # (fp64 callers get fp64 back: autograd checks the dtype)


def _aten_ctc_loss_backward(grad, log_probs, targets, input_lengths, target_lengths, neg_log_likelihood, log_alpha,
                            blank, zero_infinity=False):
    """aten::_ctc_loss_backward on CUDA; `log_alpha` is what `_aten_ctc_loss` returned."""
    if log_alpha.dtype == torch.float32:    # the unit gradient, already computed
        B = log_alpha.shape[1]
        return (log_alpha * grad.to(torch.float32).expand(B).view(1, B, 1)).to(log_probs.dtype).reshape(log_probs.shape)
    lp, tg, tgt_off, in_len, tgt_len, lmax = _prepare(log_probs, targets, list(input_lengths), list(target_lengths),
                                                       blank)
    L = _lib.lib()
    T, B, V = lp.shape
    g = grad.to(torch.float32).expand(B).contiguous()
    nll32 = neg_log_likelihood.to(torch.float32).contiguous()
    out = _launch_backward(L, False, g, lp, tg, tgt_off, in_len, tgt_len, lmax, int(blank), bool(zero_infinity),
                           nll32, log_alpha, log_alpha.numel())
    return out.to(log_probs.dtype).reshape(log_probs.shape)


def install(mode: str = "functional") -> None:
    """Route CTC-loss calls on CUDA tensors to the sm_100a kernels.

    mode="functional": replace torch.nn.functional.ctc_loss (HF, SpeechBrain and NeMo all end up there); lengths
        stay on the device, the reduction is fused.  CPU tensors keep going to torch's own CPU kernel.
    mode="aten": register CUDA implementations of aten::_ctc_loss / aten::_ctc_loss_backward through
        torch.library, below autograd: every caller (including C++ ones) is covered and ATen's composite
        ctc_loss keeps applying `reduction` / `zero_infinity`.  Cannot be undone within the process."""
    global _torch_ctc_loss, _aten_lib
    import torch.nn.functional as F
    _lib.lib()  # fail now if the library is missing
    if mode == "aten":
        if _aten_lib is None:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")   # "Overriding a previously registered kernel"
                lib_ = torch.library.Library("aten", "IMPL")
                lib_.impl("_ctc_loss", _aten_ctc_loss, "CUDA")
                lib_.impl("_ctc_loss_backward", _aten_ctc_loss_backward, "CUDA")
            _aten_lib = lib_
        return
    if mode != "functional":
        raise ValueError(mode)
    if _torch_ctc_loss is not None:
        return
    _torch_ctc_loss = F.ctc_loss

    def patched(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean",
                zero_infinity=False):
        if isinstance(log_probs, torch.Tensor) and log_probs.is_cuda and _covers(log_probs, targets, target_lengths):
            return ctc_loss(log_probs, targets, input_lengths, target_lengths, blank, reduction, zero_infinity)
        # CPU tensors, and the shapes the kernels do not cover (V > ~2040: large BPE vocabularies, target length
        # > 4095, T > 300000, an empty batch, dtypes other than fp32/fp16/bf16): torch's own implementation
        return _torch_ctc_loss(log_probs, targets, input_lengths, target_lengths, blank, reduction,
                               zero_infinity)

    patched.__wrapped__ = _torch_ctc_loss
    F.ctc_loss = patched


def uninstall() -> None:
    global _torch_ctc_loss
    import torch.nn.functional as F
    if _torch_ctc_loss is not None:
        F.ctc_loss = _torch_ctc_loss
        _torch_ctc_loss = None
