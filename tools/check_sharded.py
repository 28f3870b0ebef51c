"""torchrun --nproc-per-node N tools/check_sharded.py : utterance-sharded loss over NCCL == single-GPU loss."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import ssak_b200
from ssak_b200.shard import lattice_cost, lpt_partition, sharded_ctc_loss
from ssak_b200.synth import ctc_batch
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lp, tg, il, tl = ctc_batch(37, 300, 50, 20, 80, 7, Tmin=120)
parts = lpt_partition(lattice_cost(il.tolist(), tl.tolist()), world)
mine = parts[rank]
for red in ("mean", "sum", "mean_volume"):
    x = lp[:, mine].cuda().requires_grad_(True)
    loss = sharded_ctc_loss(x, tg[mine].cuda(), il[mine].cuda(), tl[mine].cuda(), reduction=red, zero_infinity=True, global_batch=37)
    loss.backward()
    xf = lp.cuda().requires_grad_(True)
    ref = ssak_b200.ctc_loss(xf, tg.cuda(), il.cuda(), tl.cuda(), reduction=red, zero_infinity=True)
    ref.backward()
    el = abs(loss.item() - ref.item()) / abs(ref.item())
    eg = (x.grad - xf.grad[:, mine]).abs().max().item()
    assert el < 1e-6 and eg < 1e-7, (red, el, eg)
    if rank == 0:
        print(f"sharded {red}: world={world} loss rel diff {el:.2e}, grad max diff {eg:.2e}, shard sizes {[len(p) for p in parts]}")
dist.barrier(); dist.destroy_process_group()
