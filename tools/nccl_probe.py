"""Minimal NCCL sanity check: init + one all-reduce per rank (torchrun)."""
import os, time, faulthandler
faulthandler.dump_traceback_later(45, exit=True)
import torch, torch.distributed as dist
r = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(r)
t0 = time.time()
dist.init_process_group("nccl", device_id=torch.device("cuda", r))
x = torch.ones(2, device="cuda", dtype=torch.float64) * (r + 1)
dist.all_reduce(x)
torch.cuda.synchronize()
print(f"rank {r}: all_reduce ok {x.tolist()} in {time.time()-t0:.1f}s", flush=True)
dist.barrier()
dist.destroy_process_group()
