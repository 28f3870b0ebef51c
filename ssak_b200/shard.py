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


def sharded_ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean",
                     zero_infinity=False, global_batch=None, group=None, loss_fn=None):
    """CTC loss of this rank's utterance shard, reduced over all ranks.  `loss_fn(..., reduction='none')`
    defaults to ssak_b200.ctc_loss; the CPU tests inject a stand-in."""
    import torch.distributed as dist
    if loss_fn is None:
        from .loss import ctc_loss as loss_fn
    nll = loss_fn(log_probs, targets, input_lengths, target_lengths, blank=blank, reduction="none",
                  zero_infinity=zero_infinity)
    tl = torch.as_tensor(target_lengths).to(nll.device)
    if global_batch is None:
        n = torch.tensor([nll.numel()], dtype=torch.int64, device=nll.device)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(n, group=group)
        global_batch = int(n.item())
    return reduce_loss(nll, tl, reduction, global_batch, group)
