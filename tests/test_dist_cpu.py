"""world_size-2 gloo test of the multi-GPU host logic: LPT utterance sharding + the single scalar
all-reduce reproduce the single-process loss and its gradient.  The lattice compute is a stand-in
(torch's CPU ctc_loss) because there is no GPU here; on the box the same code runs over NCCL."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_loss(lp, tg, il, tl, blank=0, reduction="none", zero_infinity=False):
    return torch.nn.functional.ctc_loss(lp, tg, il, tl, blank=blank, reduction=reduction, zero_infinity=zero_infinity)


def _worker(rank, world, port, reduction, out):
    import torch.distributed as dist
    from ssak_b200.shard import lattice_cost, lpt_partition, sharded_ctc_loss
    from ssak_b200.synth import ctc_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lp, tg, il, tl = ctc_batch(10, 40, 12, 2, 12, 5, Tmin=25, planted=False)
        mine = lpt_partition(lattice_cost(il.tolist(), tl.tolist()), world)[rank]
        x = lp[:, mine].clone().requires_grad_(True)
        loss = sharded_ctc_loss(x, tg[mine], il[mine], tl[mine], reduction=reduction, zero_infinity=True,
                                global_batch=10, loss_fn=_cpu_loss)
        loss.backward()
        # plain numpy through the queue: a torch tensor travels as a shared-memory handle that dies with the worker
        out.put((rank, mine, float(loss), x.grad.detach().numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("reduction", ["mean", "sum", "mean_volume"])
def test_sharded_loss_matches_single_process(reduction):
    from ssak_b200.synth import ctc_batch
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, reduction, q)) for r in range(2)]
    [p.start() for p in procs]
    results = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    lp, tg, il, tl = ctc_batch(10, 40, 12, 2, 12, 5, Tmin=25, planted=False)
    x = lp.clone().requires_grad_(True)
    nll = _cpu_loss(x, tg, il, tl, reduction="none", zero_infinity=True)
    ref = {"mean": (nll / tl.clamp_min(1)).mean(), "sum": nll.sum(), "mean_volume": nll.sum() / tl.sum()}[reduction]
    ref.backward()
    seen = []
    for rank, mine, loss, grad in results:
        assert abs(loss - float(ref)) <= 5e-5 * abs(float(ref)), (rank, loss, float(ref), mine)   # fp32 sums in a different order
        assert torch.allclose(torch.from_numpy(grad), x.grad[:, mine], atol=1e-6)
        seen += mine
    assert sorted(seen) == list(range(10))


def _worker_empty(rank, world, port, out):
    import torch.distributed as dist
    from ssak_b200.shard import lattice_cost, lpt_partition, sharded_ctc_loss
    from ssak_b200.synth import ctc_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lp, tg, il, tl = ctc_batch(2, 30, 9, 2, 8, 6, Tmin=20, planted=False)
        mine = lpt_partition(lattice_cost(il.tolist(), tl.tolist()), world)[rank]   # 2 utterances, 3 ranks: one is empty
        x = lp[:, mine].clone().requires_grad_(True)
        loss = sharded_ctc_loss(x, tg[mine], il[mine], tl[mine], reduction="mean", zero_infinity=True,
                                global_batch=2, loss_fn=_cpu_loss)
        loss.backward()
        out.put((rank, mine, float(loss), x.grad.detach().numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_empty_shard_takes_part_in_the_collective():
    """Fewer utterances than ranks: the rank with an empty shard must not hang the others in the all-reduce; it
    gets the global loss and an empty gradient."""
    from ssak_b200.synth import ctc_batch
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_empty, args=(r, 3, port, q)) for r in range(3)]
    [p.start() for p in procs]
    results = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    lp, tg, il, tl = ctc_batch(2, 30, 9, 2, 8, 6, Tmin=20, planted=False)
    ref = float((_cpu_loss(lp, tg, il, tl, reduction="none", zero_infinity=True) / tl.clamp_min(1)).mean())
    assert sorted(len(m) for _, m, _, _ in results) == [0, 1, 1]
    for rank, mine, loss, grad in results:
        assert abs(loss - ref) <= 5e-5 * abs(ref), (rank, loss, ref)
        assert grad.shape[1] == len(mine)
