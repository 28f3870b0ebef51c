"""Utterance sharding across the GPUs of one box.

Utterances are independent, so the batch is partitioned by utterance with a
longest-processing-time-first assignment on the lattice cost T_b*(2L_b+1) (loss) or T_b*(L_b+1)
(alignment); no lattice data crosses GPUs.  The only exchange of the loss is ONE all-reduce of
two scalars per step, [sum_b nll_b*w_b, sum_b weight_b], over NCCL (NVLink 5 / NVSwitch); gloo
runs the same code on CPU tensors in the tests.  Gradients w.r.t. the local emissions need no
exchange: the 'mean' scaling only needs the global batch size, known a priori.
"""
from __future__ import annotations

from typing import List, Sequence

import torch


def lattice_cost(input_lengths: Sequence[int], target_lengths: Sequence[int], kind: str = "loss") -> List[int]:
    mul = (lambda l: 2 * l + 1) if kind == "loss" else (lambda l: l + 1)
    return [int(t) * mul(int(l)) for t, l in zip(input_lengths, target_lengths)]


def lpt_partition(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first greedy: indices per rank, loads within 4/3 of optimal.
    Deterministic (ties -> lower index, lower rank), so every rank computes the same partition."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += costs[i]
    return [sorted(p) for p in parts]


def length_buckets(costs: Sequence[int], n_buckets: int) -> List[List[int]]:
    """Sort by cost and cut into n_buckets contiguous groups (similar lengths launch together, so
    CTAs of one launch finish together)."""
    order = sorted(range(len(costs)), key=lambda i: (costs[i], i))
    n = len(order)
    return [order[(k * n) // n_buckets: ((k + 1) * n) // n_buckets] for k in range(n_buckets)]


def reduce_loss(local_nll: torch.Tensor, local_target_lengths: torch.Tensor, reduction: str,
                global_batch: int, group=None) -> torch.Tensor:
    """Global reduction of per-utterance nll held by each rank: one all-reduce of 2 scalars.

    'mean': mean_b(nll_b / clamp(L_b,1)) over the GLOBAL batch; 'sum'; 'mean_volume': sum nll / sum L.
    Differentiable w.r.t. local_nll (the all-reduce result enters as sum of local + detached remote)."""
    import torch.distributed as dist
    tl = local_target_lengths.to(local_nll.dtype)
    if reduction == "mean":
        num_local, den_local = (local_nll / tl.clamp_min(1)).sum(), local_nll.new_tensor(float(local_nll.numel()))
    elif reduction == "sum":
        num_local, den_local = local_nll.sum(), local_nll.new_tensor(1.0)
    elif reduction == "mean_volume":
        num_local, den_local = local_nll.sum(), tl.sum()
    else:
        raise ValueError(reduction)
    packed = torch.stack([num_local.detach(), den_local.detach()]).to(torch.float64)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    num_glob, den_glob = packed[0].to(local_nll.dtype), packed[1].to(local_nll.dtype)
    if reduction == "mean":
        den = local_nll.new_tensor(float(global_batch))
    elif reduction == "sum":
        den = local_nll.new_tensor(1.0)
    else:
        den = den_glob.clamp_min(1)
    # value = global numerator / denominator; gradient flows through the local numerator only
    return (num_glob + (num_local - num_local.detach())) / den


_SIDE_STREAMS = {}
_SCALARS = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return st


def _device_scalar(value: float, dev):
    key = (float(value), dev.type, dev.index)
    t = _SCALARS.get(key)
    if t is None:
        if len(_SCALARS) > 256:
            _SCALARS.clear()
        t = _SCALARS[key] = torch.full((1,), float(value), dtype=torch.float32, device=dev)
    return t


class _ShardedCTCFunction(torch.autograd.Function):
    """The sharded loss on the sm_100a kernels in as few launches as the single-GPU wrapper plus the collective:
    lattice forward + join, the fused local reduction, ONE all-reduce of [numerator, denominator], one division;
    backward: one scale of the per-utterance weights and the lattice backward.

    The collective is OFF the critical path (`overlap`): for 'mean' and 'sum' the backward only needs the a-priori
    known global batch, so the all-reduce and the division run on a side stream next to the backward
    kernels and are joined to the caller's stream at the end of backward().  Until then the returned loss is
    ordered on the side stream only (it is pre-filled with NaN on the caller's stream, so a premature read is loud,
    never a stale number); without gradients, and for 'mean_volume' (whose denominator IS the collective's result),
    the join happens before forward returns.

    Where the throughput kernels run (ssak_ctc_loss_nll_is_provisional) the likelihoods are final only after the
    backward call, so forward() runs it too -- with the a-priori denominator -- packs the local sums from the final
    likelihoods and only then starts the collective; backward() applies autograd's upstream gradient (a launch
    without memory traffic when it is 1) and joins the side stream."""

    @staticmethod
    def forward(ctx, log_probs, targets, tgt_off, in_len, tgt_len, max_target_len, blank, zero_infinity,
                reduction, global_batch, group, overlap):
        import torch.distributed as dist
        from . import _lib
        from .loss import _require_supported
        L = _lib.lib()
        T, B, V = log_probs.shape
        dev = log_probs.device
        _require_supported(T, B, V, max_target_len)
        ws_bytes = L.ssak_ctc_loss_workspace_bytes_v(T, B, V, max_target_len, 1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        gscale = torch.empty(B, dtype=torch.float32, device=dev)
        packed = torch.empty(2, dtype=torch.float64, device=dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)      # [loss, 1/denominator]
        main = torch.cuda.current_stream(dev)
        stream = main.cuda_stream
        rc = L.ssak_ctc_loss_forward(log_probs.data_ptr(), T, B, V, log_probs.stride(0), log_probs.stride(1),
                                     targets.data_ptr(), tgt_off.data_ptr(), in_len.data_ptr(), tgt_len.data_ptr(),
                                     max_target_len, blank, 1, nll.data_ptr(), ws.data_ptr(), ws_bytes, stream)
        _lib.check(rc, "ssak_ctc_loss_forward")
        code = {"mean": 1, "sum": 2, "mean_volume": 3}[reduction]
        rc = L.ssak_ctc_shard_pack(nll.data_ptr(), tgt_len.data_ptr(), B, code, int(zero_infinity),
                                   packed.data_ptr(), gscale.data_ptr(), stream)
        _lib.check(rc, "ssak_ctc_shard_pack")
        distributed = dist.is_available() and dist.is_initialized()
        side = None
        # Throughput kernels (batches that fill the GPU): the likelihood is provisional until the backward call has
        # verified it (loss._CTCLossFunction), so the backward runs right here -- with the a-priori denominator for
        # 'mean' / 'sum', 1 for 'mean_volume' -- and the local sums are packed again from the final likelihoods.
        eager = bool(ctx.needs_input_grad[0]) and bool(L.ssak_ctc_loss_nll_is_provisional(B, V, max_target_len))
        ctx.eager = eager
        if eager:
            from .loss import _launch_backward
            inv0 = 1.0 / float(global_batch) if code == 1 else 1.0
            g = torch.empty(B, dtype=torch.float32, device=dev)
            rc = L.ssak_ctc_shard_grad_scale(gscale.data_ptr(), _device_scalar(1.0, dev).data_ptr(),
                                             _device_scalar(inv0, dev).data_ptr(), B, g.data_ptr(), stream)
            _lib.check(rc, "ssak_ctc_shard_grad_scale")
            grad = _launch_backward(L, False, g, log_probs, targets, tgt_off, in_len, tgt_len, max_target_len, blank,
                                    zero_infinity, nll, ws, ws_bytes)
            rc = L.ssak_ctc_shard_pack(nll.data_ptr(), tgt_len.data_ptr(), B, code, int(zero_infinity),
                                       packed.data_ptr(), gscale.data_ptr(), stream)
            _lib.check(rc, "ssak_ctc_shard_pack")
        if distributed and overlap and code != 3 and ctx.needs_input_grad[0]:
            out.fill_(float("nan"))
            side = _side_stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
                rc = L.ssak_ctc_shard_finish(packed.data_ptr(), code, int(global_batch), out.data_ptr(),
                                             out.data_ptr() + 4, side.cuda_stream)
            packed.record_stream(side)
            out.record_stream(side)
        else:
            if distributed:
                dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
            rc = L.ssak_ctc_shard_finish(packed.data_ptr(), code, int(global_batch), out.data_ptr(),
                                         out.data_ptr() + 4, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "ssak_ctc_shard_finish")
        if eager:
            ctx.save_for_backward(grad, out)
        else:
            ctx.save_for_backward(log_probs, targets, tgt_off, in_len, tgt_len, nll, ws, gscale, out)
        ctx.meta = (max_target_len, blank, zero_infinity, ws_bytes, code, int(global_batch))
        ctx.side = side
        return out[0]

    @staticmethod
    def backward(ctx, grad_loss):
        from . import _lib
        L = _lib.lib()
        max_target_len, blank, zero_infinity, ws_bytes, code, global_batch = ctx.meta
        gl = grad_loss if (grad_loss.dtype == torch.float32 and grad_loss.is_contiguous()) else grad_loss.float().contiguous()
        if ctx.eager:
            # apply autograd's upstream gradient (and, for 'mean_volume', the collective's 1/denominator); no memory
            # traffic when the factor is 1.  (A second backward over a retained graph would scale twice: not supported.)
            grad, out = ctx.saved_tensors
            T, B, V = grad.shape
            with torch.cuda.device(grad.device):
                main = torch.cuda.current_stream()
                if ctx.side is not None:
                    main.wait_stream(ctx.side)
                rc = L.ssak_ctc_grad_scale(grad.data_ptr(), T, B, V, grad.stride(0), grad.stride(1), None, gl.data_ptr(),
                                           out.data_ptr() + 4 if code == 3 else None, main.cuda_stream)
            _lib.check(rc, "ssak_ctc_grad_scale")
            return (grad,) + (None,) * 11
        log_probs, targets, tgt_off, in_len, tgt_len, nll, ws, gscale, out = ctx.saved_tensors
        T, B, V = log_probs.shape
        dev = log_probs.device
        g = torch.empty(B, dtype=torch.float32, device=dev)
        grad = torch.empty_like(log_probs)
        if grad.stride(2) != 1:
            grad = torch.empty((T, B, V), dtype=torch.float32, device=dev)
        # 1/denominator: known a priori for 'mean' (global batch) and 'sum' (1); only 'mean_volume' takes it from
        # the collective's result
        if code == 1:
            inv_den_ptr = _device_scalar(1.0 / float(global_batch), dev).data_ptr()
        elif code == 2:
            inv_den_ptr = _device_scalar(1.0, dev).data_ptr()
        else:
            inv_den_ptr = out.data_ptr() + 4
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            stream = main.cuda_stream
            rc = L.ssak_ctc_shard_grad_scale(gscale.data_ptr(), gl.data_ptr(), inv_den_ptr, B, g.data_ptr(), stream)
            _lib.check(rc, "ssak_ctc_shard_grad_scale")
            rc = L.ssak_ctc_loss_backward(g.data_ptr(), log_probs.data_ptr(), T, B, V, log_probs.stride(0),
                                          log_probs.stride(1), targets.data_ptr(), tgt_off.data_ptr(),
                                          in_len.data_ptr(), tgt_len.data_ptr(), max_target_len, blank,
                                          int(zero_infinity), nll.data_ptr(), grad.data_ptr(), grad.stride(0),
                                          grad.stride(1), ws.data_ptr(), ws_bytes, stream)
            if ctx.side is not None:     # end of the step: the loss value becomes ordered on the caller's stream
                main.wait_stream(ctx.side)
        _lib.check(rc, "ssak_ctc_loss_backward")
        return (grad,) + (None,) * 11


def _empty_shard_loss(log_probs, reduction, global_batch, group):
    """A rank whose shard is empty (fewer utterances than ranks) still takes part in the collective with zeros and
    returns the global loss, connected to the graph with a zero gradient."""
    import torch.distributed as dist
    packed = torch.zeros(2, dtype=torch.float64, device=log_probs.device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    if reduction == "mean":
        den = float(global_batch)
    elif reduction == "sum":
        den = 1.0
    else:
        den = packed[1].clamp_min(1.0)
    return (packed[0] / den).to(torch.float32) + 0.0 * log_probs.sum()


def sharded_ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean",
                     zero_infinity=False, global_batch=None, group=None, loss_fn=None, overlap=True):
    """CTC loss of this rank's utterance shard, reduced over all ranks.

    Default: the sm_100a kernels through one lean autograd node (`_ShardedCTCFunction`).  With `loss_fn`
    (`loss_fn(..., reduction='none')`, e.g. the CPU stand-in of the gloo tests) the same value is built from
    generic torch ops.

    The value is the loss over the GLOBAL batch and the gradient is d(global loss)/d(local log_probs).  Under
    DistributedDataParallel the parameter gradients are additionally AVERAGED over the ranks, so the effective
    gradient is 1/world_size of the single-GPU one: multiply the loss by the world size (or use reduction='sum' and
    normalise yourself) to reproduce the single-process numbers.  `overlap` (default): the all-reduce runs beside
    the backward kernels, see `_ShardedCTCFunction`; read the loss value after backward(), or pass overlap=False."""
    import torch.distributed as dist
    if reduction not in ("mean", "sum", "mean_volume"):
        raise ValueError(reduction)
    if loss_fn is None:
        if log_probs.dim() == 3 and log_probs.shape[1] == 0:
            if global_batch is None:
                n = torch.zeros(1, dtype=torch.int64, device=log_probs.device)
                if dist.is_available() and dist.is_initialized():
                    dist.all_reduce(n, group=group)
                global_batch = max(int(n.item()), 1)
            return _empty_shard_loss(log_probs, reduction, global_batch, group)
        from .loss import _prepare
        lp, tg, tgt_off, in_len, tgt_len, lmax = _prepare(log_probs, targets, input_lengths, target_lengths, blank)
        if global_batch is None:
            n = torch.tensor([lp.shape[1]], dtype=torch.int64, device=lp.device)
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(n, group=group)
            global_batch = int(n.item())
        with torch.cuda.device(lp.device):
            return _ShardedCTCFunction.apply(lp, tg, tgt_off, in_len, tgt_len, lmax, int(blank), bool(zero_infinity),
                                             reduction, int(global_batch), group, bool(overlap))
    if log_probs.dim() == 3 and log_probs.shape[1] == 0:
        if global_batch is None:
            n = torch.zeros(1, dtype=torch.int64, device=log_probs.device)
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(n, group=group)
            global_batch = max(int(n.item()), 1)
        return _empty_shard_loss(log_probs, reduction, global_batch, group)
    nll = loss_fn(log_probs, targets, input_lengths, target_lengths, blank=blank, reduction="none",
                  zero_infinity=zero_infinity)
    tl = torch.as_tensor(target_lengths).to(nll.device)
    if global_batch is None:
        n = torch.tensor([nll.numel()], dtype=torch.int64, device=nll.device)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(n, group=group)
        global_batch = int(n.item())
    return reduce_loss(nll, tl, reduction, global_batch, group)
