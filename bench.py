#!/usr/bin/env python
"""bench.py -- lattice-cells/s of the CTC lattice hot path (loss fwd+bwd, forced alignment, greedy) on B200.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path, rank 0)

Headline workload (the configuration BASELINE.json's north star quotes its target on): the "1k-utterance batch" --
synthetic CTC loss forward + backward, B=1024 utterances per GPU, T=1500 frames, V=50, L in [200,400],
T_b in [1200,1500], fp32, planted-alignment emissions.  A step = one forward + one
backward of the loss over the batch; cells = sum_b T_b*(2L_b+1), counted once per fwd+bwd.  Rank 0 prints ONE JSON
line.  At N=1 the line also carries, under "extra", every other configuration of BASELINE.json measured in the same
process (C2 loss, C4, C5 loss / logits path, forced alignment C2-shape / C5 / C3 with an end-to-end leg through
ssak_forced_align_host, greedy, and torch's own CUDA ctc_loss as the on-box GPU comparator), each with its own
roofline block.  --workload c2|c4|c5 selects another headline; --scaling strong splits ONE global batch of the
workload over the ranks (LPT partition on the lattice cost + length buckets) instead of one batch per rank.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B, T, V, Lmin, Lmax, Tmin)
    "c2": (64, 1500, 50, 200, 400, 1200),
    "1k": (1024, 1500, 50, 200, 400, 1200),
    "c4": (256, 1500, 50, 81, 405, 300),      # ragged: T_b ~ U[300,1500], L_b = 0.27 T_b
    "c5": (512, 750, 1024, 100, 200, 600),
}
METRIC = "ctc_lattice_cells_per_sec"
UNIT = "cells/s"


def c4_batch(seed=1234 + 4):
    """BASELINE config C4: 256 utterances, T_b ~ U[300,1500], L_b = 0.27 T_b (13 characters per second), V = 50."""
    from ssak_b200.synth import planted_emissions
    g = torch.Generator().manual_seed(seed)
    B, T, V = 256, 1500, 50
    il = torch.randint(300, T + 1, (B,), generator=g)
    tl = (0.27 * il.float()).round().long().clamp_min(1)
    tg = torch.randint(1, V, (B, int(tl.max())), generator=g)
    lp = torch.empty(T, B, V)
    for b in range(B):
        e = torch.randn(T, V, generator=g)
        e[: int(il[b])] = planted_emissions(int(il[b]), V, tg[b, : int(tl[b])], g, 0, normalize=False)
        lp[:, b] = e.log_softmax(-1)
    return lp, tg, il, tl


def make_batch(name, seed):
    from ssak_b200.synth import ctc_batch
    if name == "c4":
        lp, tg, il, tl = c4_batch(seed)
    else:
        B, T, V, Lmin, Lmax, Tmin = WORKLOADS[name]
        lp, tg, il, tl = ctc_batch(B, T, V, Lmin, Lmax, seed, Tmin=Tmin, planted=True)
    cells = int((il * (2 * tl + 1)).sum())
    return lp, tg, il, tl, cells


def describe(name):
    B, T, V, Lmin, Lmax, Tmin = WORKLOADS[name]
    if name == "c4":
        return f"c4: CTC loss fwd+bwd, B={B} ragged (T_b ~ U[300,{T}], L_b = 0.27 T_b), V={V}"
    tag = "1k-utterance batch" if name == "1k" else name
    return (f"{tag}: CTC loss fwd+bwd, B={B}, T={T}, V={V}, L in [{Lmin},{Lmax}], T_b in [{Tmin},{T}], "
            f"planted-alignment emissions, reduction=mean, zero_infinity")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under the benchmark's load (one background
    `nvidia-smi -lms 20` process, as in B200_PROFILING.md; started before the timed region, stopped after)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.samples = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            first = self.proc.stdout.readline()   # nvidia-smi needs ~1 s to start: wait for its first sample
            f = [x.strip() for x in first.split(",")]
            if len(f) >= 6 and f[0].replace(".", "").isdigit():
                self.samples.append(f)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6 and f[0].replace(".", "").isdigit():
                self.samples.append(f)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        busy = [x for x in sm if x > 0.5 * sm[-1]] or sm   # samples taken while the kernels were running
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples), "window": "timed steps + 0.3 s of the same step (soak) + e2e loop"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_of(key):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        return json.load(open(tp)).get(key)
    return None


def cpu_reference_time(lp, tg, il, tl, repeats=3):
    """The reference's CPU path for the loss: torch CPU F.ctc_loss fwd + bwd, all host threads."""
    import torch.nn.functional as F
    best = float("inf")
    for _ in range(repeats):
        x = lp.clone().requires_grad_(True)
        t0 = time.perf_counter()
        F.ctc_loss(x, tg, il, tl, blank=0, reduction="mean", zero_infinity=True).backward()
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_sample(name, lp, tg, il, tl):
    """A bounded sample of the workload for the CPU legs (the full 1k batch takes ~3 s per step on 16 threads)."""
    n = {"1k": 128, "c5": 64, "c4": 128, "c2": 64}[name]
    n = min(n, lp.shape[1])
    cells = int((il[:n] * (2 * tl[:n] + 1)).sum())
    return lp[:, :n].contiguous(), tg[:n], il[:n], tl[:n], cells, n


def run_reference(args, rank):
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    torch.set_num_threads(ncores)
    name = args.workload
    lp, tg, il, tl, _ = make_batch(name, 1234 + 2)
    lp, tg, il, tl, cells, n = cpu_sample(name, lp, tg, il, tl)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_time(lp, tg, il, tl, 1)
    times = [cpu_reference_time(lp, tg, il, tl, 1) for _ in range(args.steps)]
    t = sum(times) / len(times)
    val = cells / t
    sample = (f"the first {n} utterances of the {name} batch per step, torch {torch.__version__} CPU F.ctc_loss fwd+bwd "
              f"(the arithmetic the reference reaches), {ncores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": describe(name), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def time_kernels(lib, dev, lp_d, tg32, off, il32, tl32, Lmax, iters, flush, logits=False):
    """CUDA-event time of the forward launch group and of the backward launch group, separately."""
    T, B, V = lp_d.shape
    ws_bytes = lib.ssak_ctc_loss_workspace_bytes_v(T, B, V, Lmax, 1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    nll = torch.empty(B, dtype=torch.float32, device=dev)
    go = torch.full((B,), 1.0 / B, dtype=torch.float32, device=dev)
    grad = torch.empty_like(lp_d)
    s = torch.cuda.current_stream().cuda_stream
    fwd = lib.ssak_ctc_logits_forward if logits else lib.ssak_ctc_loss_forward
    bwd = lib.ssak_ctc_logits_backward if logits else lib.ssak_ctc_loss_backward
    tf, tb = [], []
    for i in range(iters + 2):
        flush.add_(1)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        rc = fwd(lp_d.data_ptr(), T, B, V, lp_d.stride(0), lp_d.stride(1), tg32.data_ptr(), off.data_ptr(),
                 il32.data_ptr(), tl32.data_ptr(), Lmax, 0, 1, nll.data_ptr(), ws.data_ptr(), ws_bytes, s)
        assert rc == 0, rc
        e1.record()
        rc = bwd(go.data_ptr(), lp_d.data_ptr(), T, B, V, lp_d.stride(0), lp_d.stride(1), tg32.data_ptr(),
                 off.data_ptr(), il32.data_ptr(), tl32.data_ptr(), Lmax, 0, 1, nll.data_ptr(), grad.data_ptr(),
                 grad.stride(0), grad.stride(1), ws.data_ptr(), ws_bytes, s)
        assert rc == 0, rc
        e2.record()
        torch.cuda.synchronize()
        if i >= 2:
            tf.append(e0.elapsed_time(e1))
            tb.append(e1.elapsed_time(e2))
    return statistics.mean(tf) * 1e-3, statistics.mean(tb) * 1e-3, ws_bytes


def loss_roofline(name, V, sumT, t_fwd, t_bwd, hbm, hbm_src, ws_bytes=None):
    """Roofline block of the loss: the dominant launch is the backward (recursion + gradient): it reads every
    emission row once and writes the gradient row (8 B per (t,b,v)); the forward reads the rows once (4 B)."""
    sumTV = float(sumT) * V
    bwd_bytes = 8.0 * sumTV
    achieved = bwd_bytes / t_bwd / 1e9
    return {"bound": "hbm", "kernel": "loss backward launch (recursion + posteriors + gradient rows)",
            "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
            "traffic": traffic_of(name), "peak_source": hbm_src,
            "algorithmic_bytes_per_launch": bwd_bytes, "launch_ms": t_bwd * 1e3,
            "forward_launch_ms": t_fwd * 1e3, "forward_algorithmic_bytes": 4.0 * sumTV,
            "forward_achieved_gbs": 4.0 * sumTV / t_fwd / 1e9, "forward_frac": 4.0 * sumTV / t_fwd / 1e9 / hbm,
            "step_frac_12B": 12.0 * sumTV / (t_fwd + t_bwd) / 1e9 / hbm,
            "workspace_bytes": ws_bytes,
            "note": "per cell the compulsory traffic is 0.75 B at V=50 and 30.6 B at V=1024: the V=50 shapes are "
                    "bound by the recursion's instruction issue (secondary bound, DESIGN.md), V=1024 by HBM"}


def _set_affinity(local_rank, world):
    """Give every rank its own slice of the host cores (the e2e leg is host-memory / PCIe work: 8 ranks on the same
    cores of NUMA node 0 cost 2.5x at N=8 in round 1).  The slice follows the GPU's NUMA node when sysfs tells."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        node_cores = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            path = f"/sys/bus/pci/devices/{bus.lower()[-12:]}/local_cpulist"
            if os.path.exists(path):
                node_cores = []
                for part in open(path).read().strip().split(","):
                    a, _, b = part.partition("-")
                    node_cores += list(range(int(a), int(b or a) + 1))
                node_cores = [c for c in node_cores if c in cores]
        except Exception:
            node_cores = None
        pool = node_cores or cores
        per = max(1, len(pool) // max(world, 1))
        mine = pool[(local_rank * per) % len(pool):][:per] or pool
        os.sched_setaffinity(0, mine)
        return {"cores": len(mine), "first": mine[0], "numa_local": bool(node_cores)}
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import ssak_b200
    from ssak_b200.shard import lattice_cost, length_buckets, lpt_partition, sharded_ctc_loss
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    aff = _set_affinity(local_rank, world) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = ssak_b200.lib()
    name = args.workload
    B, T, V, Lmin, Lmax, Tmin = WORKLOADS[name]
    strong = args.scaling == "strong"
    if strong:
        # ONE global batch, split by the longest-processing-time-first partition on the lattice cost; every rank
        # generates the same batch (same seed) and keeps its share
        lp, tg, il, tl, _ = make_batch(name, 1234 + 2)
        parts = lpt_partition(lattice_cost(il.tolist(), tl.tolist()), world)
        mine = parts[rank]
        lp, tg, il, tl = lp[:, mine].contiguous(), tg[mine], il[mine], tl[mine]
        cells = int((il * (2 * tl + 1)).sum())
        global_batch = B
    else:
        lp, tg, il, tl, cells = make_batch(name, 1234 + 2 + 1000 * rank)   # weak scaling: own batch per rank
        global_batch = B * world
    Bl = lp.shape[1]
    # length buckets (similar lengths launch together): only worth it for ragged batches with enough utterances
    nb = 2 if (name == "c4" and Bl >= 64) else 1
    buckets = [sorted(bk) for bk in length_buckets(lattice_cost(il.tolist(), tl.tolist()), nb)] if nb > 1 else [list(range(Bl))]
    lp_pin, grad_pin = lp.pin_memory(), torch.empty_like(lp).pin_memory()
    lp_d = lp_pin.to(dev, non_blocking=True)
    dev_b = []
    for bk in buckets:
        idx = torch.as_tensor(bk)
        Tm = int(il[idx].max())
        x = lp_d[:Tm, idx.to(dev)].contiguous() if nb > 1 else lp_d
        dev_b.append((x, tg[idx].to(dev, torch.int32), il[idx].to(dev, torch.int32), tl[idx].to(dev, torch.int32)))
    flush = torch.zeros(96 * 1024 * 1024, dtype=torch.float32, device=dev)      # 384 MB > 126 MB L2

    def step():
        grads, loss = [], None
        for x0, tg_d, il_d, tl_d in dev_b:
            x = x0.detach().requires_grad_(True)
            if world > 1 or nb > 1:   # utterance-sharded: local lattices + ONE all-reduce of 2 scalars per call (NCCL),
                                      # off the critical path (side stream, joined at the end of backward)
                l = sharded_ctc_loss(x, tg_d, il_d, tl_d, blank=0, reduction="mean", zero_infinity=True,
                                     global_batch=global_batch)
            else:
                l = ssak_b200.ctc_loss(x, tg_d, il_d, tl_d, blank=0, reduction="mean", zero_infinity=True)
            l.backward()
            grads.append(x.grad)
            loss = l if loss is None else loss + l
        return loss, grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident throughput: K steps, CUDA events on the launching stream, L2 flushed
    #      (384 MB write) before each step, outside the per-step event pair
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        evs.append((a, b))
    barrier()
    t_dev = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
    # the timed region can be shorter than nvidia-smi's sampling period: keep the same work running for ~0.3 s more
    # so that the clock / throttle samples are taken under this load (every rank runs the same fixed number of
    # steps: the sharded step contains a collective)
    t_soak = torch.tensor([t_dev / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_soak, op=dist.ReduceOp.MAX)       # (a count derived from the local clock differs between ranks)
    n_soak = max(2, min(600, int(0.3 / max(t_soak.item(), 1e-4))))
    for _ in range(n_soak):
        step()
    torch.cuda.synchronize()
    # ---- end to end through the host-buffer C ABI: pinned host log-probs in, nll + gradient out
    ctx = C.c_void_p()
    assert lib.ssak_context_create(local_rank, C.byref(ctx)) == 0
    tg32 = tg.to(torch.int32).contiguous()
    il32, tl32 = il.to(torch.int32), tl.to(torch.int32)
    nll_h = torch.empty(Bl, dtype=torch.float32).pin_memory()

    def e2e_step():
        rc = lib.ssak_ctc_loss_host(ctx, lp_pin.data_ptr(), lp.shape[0], Bl, V, tg32.data_ptr(), tg32.shape[1],
                                    il32.data_ptr(), tl32.data_ptr(), 0, 1, None, nll_h.data_ptr(),
                                    grad_pin.data_ptr())
        assert rc == 0, rc

    n_e2e = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    t_e2e = (time.perf_counter() - t0) / n_e2e * args.steps     # scaled to K steps (same aggregate below)
    if rank == 0:
        sampler.stop()
    # parity gate on the timed configuration: the host-ABI result equals the torch-facing one
    if nb == 1:
        _, g = step()
        torch.cuda.synchronize()
        gscale = (1.0 / (global_batch * tl.clamp_min(1).float())).view(1, Bl, 1)
        assert (g[0].cpu() - grad_pin * gscale).abs().max().item() < 1e-6

    times = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
    tot_cells = torch.tensor([float(cells)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_cells, op=dist.ReduceOp.SUM)
    t_dev, t_e2e = times.tolist()
    total = tot_cells.item()

    out = None
    if rank == 0:
        hbm, hbm_src = peaks()
        off = (torch.arange(Bl, device=dev, dtype=torch.int64) * tg.shape[1])
        t_fwd, t_bwd, ws_bytes = time_kernels(lib, dev, lp_d, tg32.to(dev), off, il32.to(dev), tl32.to(dev),
                                              int(tl.max()), max(min(args.steps, 10), 5), flush)
        h2d = lp.numel() * 4 + tg32.numel() * 4 + 2 * Bl * 4 + Bl * 8 + Bl * 4
        d2h = lp.numel() * 4 + Bl * 4
        out = {
            "metric": METRIC, "value": total * args.steps / t_dev, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": describe(name) + (f"; ONE global batch of {B} split over {world} GPUs (LPT on "
                                                     f"T_b(2L_b+1)), {nb} length bucket(s) per rank" if strong
                                                     else f"; B={B} per GPU"),
                       "cells_per_step_rank0": cells, "utterances_rank0": Bl,
                       "l2": "flushed (384 MB write) before every timed step",
                       "inputs": "log_probs fp32 [T,B,V] contiguous, int32 padded targets / lengths, resident in HBM",
                       "timing": "per-step CUDA events on the launching stream, summed; max over ranks",
                       "collective": None if world == 1 else "one ncclAllReduce of 2 doubles per loss call, on a "
                                                             "side stream, joined at the end of backward",
                       "host_affinity": aff},
            "e2e": {"value": total * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e / args.steps * 1e3, "steps_timed": n_e2e,
                    "path": "ssak_ctc_loss_host (C ABI): pinned host log-probs in, nll + full gradient out"},
            "gpu_launches": None,
            "roofline": loss_roofline(name, V, int(il.sum()), t_fwd, t_bwd, hbm, hbm_src, ws_bytes),
            "clocks": sampler.summary(),
        }
        out["gpu_launches"] = sum(launches_per_step(lib, x.shape[1], V, int(t.max())) for x, _, _, t in dev_b) * args.steps
        if world == 1:
            torch.set_num_threads(os.cpu_count() or 1)
            slp, stg, sil, stl, scells, sn = cpu_sample(name, lp, tg, il, tl)
            cpu_s = cpu_reference_time(slp, stg, sil, stl, 3)
            out["cpu_baseline"] = {"value": scells / cpu_s, "unit": UNIT, "cores": torch.get_num_threads(),
                                   "kind": "reference",
                                   "sample": f"the first {sn} utterances of the batch, fwd+bwd, best of 3, torch "
                                             f"{torch.__version__} CPU F.ctc_loss with {torch.get_num_threads()} threads"}
            if not args.no_extra:
                del lp_d, dev_b
                torch.cuda.empty_cache()
                out["extra"] = extra_numbers(lib, dev, flush, skip=name)
    lib.ssak_context_destroy(ctx)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def launches_per_step(lib, B, V, Lmax):
    """Kernels of ours per loss call (forward group + backward group + the reduction), as the library reports them;
    where the likelihood is final only after the backward call (throughput kernels) the wrapper reduces twice and
    launches the upstream-gradient kernel."""
    n = int(lib.ssak_ctc_loss_launches(B, V, Lmax, 0)) + 1
    if lib.ssak_ctc_loss_nll_is_provisional(B, V, Lmax):
        n += 2
    return n


def _timed_ms(fn, flush, n=7, warm=3):
    """median device time of fn() over n calls (CUDA events, L2 flushed before each), after `warm` untimed calls
    (the first calls of a kernel pay its lazy module load)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def extra_numbers(lib, dev, flush, skip=None):
    """The other configurations of BASELINE.json in the same process (rank 0, N = 1): device-resident kernels,
    each with its roofline block; the aligner also end to end through the host-buffer C ABI."""
    import torch.nn.functional as F
    import ssak_b200
    from ssak_b200.synth import align_batch
    hbm, hbm_src = peaks()
    res = {}
    batches = {}
    # ---- loss: C2 / 1k / C5 / C4, kernels only (forward launch group, backward launch)
    for name in ("c2", "1k", "c5", "c4"):
        try:
            B, T, V, Lmin, Lmax, Tmin = WORKLOADS[name]
            lp, tg, il, tl, cells = make_batch(name, 99 if name != "c4" else 1234 + 4)
            lp_d = lp.to(dev)
            off = torch.arange(lp.shape[1], device=dev, dtype=torch.int64) * tg.shape[1]
            tf, tb, wsb = time_kernels(lib, dev, lp_d, tg.to(torch.int32).to(dev), off, il.to(torch.int32).to(dev),
                                       tl.to(torch.int32).to(dev), int(tl.max()), 5, flush)
            res[f"loss_{name}"] = {"cells_per_s": cells / (tf + tb), "fwd_ms": tf * 1e3, "bwd_ms": tb * 1e3,
                                   "workload": describe(name),
                                   "roofline": loss_roofline(name, V, int(il.sum()), tf, tb, hbm, hbm_src, wsb)}
            if name != "c4":
                batches[name] = (lp_d, tg, il, tl, cells)
            if name == "c4":
                # the same batch as 4 length buckets (what a bucketed sampler hands over; ssak_b200.shard)
                from ssak_b200.shard import lattice_cost, length_buckets
                t_bkt = 0.0
                for bk in length_buckets(lattice_cost(il.tolist(), tl.tolist()), 4):
                    idx = torch.as_tensor(sorted(bk))
                    Tm = int(il[idx].max())
                    sub = lp[:Tm, idx].contiguous().to(dev)
                    o2 = torch.arange(len(idx), device=dev, dtype=torch.int64) * tg.shape[1]
                    a, b2, _ = time_kernels(lib, dev, sub, tg[idx].to(torch.int32).to(dev), o2,
                                            il[idx].to(torch.int32).to(dev), tl[idx].to(torch.int32).to(dev),
                                            int(tl[idx].max()), 5, flush)
                    t_bkt += a + b2
                res["loss_c4"]["ms_4_length_buckets"] = t_bkt * 1e3
                res["loss_c4"]["cells_per_s_4_length_buckets"] = cells / t_bkt
            del lp
        except Exception as e:  # keep the headline line even if an extra shape fails
            res[f"loss_{name}"] = {"error": repr(e)}
    # ---- on-box GPU comparator: torch's own CUDA ctc_loss (native kernel, cuDNN off as HF does) on the
    #      same tensors, forward + backward through autograd, same event timing and L2 flush
    for name in ("c2", "1k", "c5"):
        if name not in batches:
            continue
        try:
            lp_d, tg, il, tl, cells = batches[name]
            tg_d, il_d, tl_d = tg.to(dev), il.to(dev), tl.to(dev)

            def torch_step():
                x = lp_d.detach().requires_grad_(True)
                with torch.backends.cudnn.flags(enabled=False):
                    F.ctc_loss(x, tg_d, il_d, tl_d, 0, "mean", True).backward()

            def our_step():
                x = lp_d.detach().requires_grad_(True)
                ssak_b200.ctc_loss(x, tg_d, il_d, tl_d, 0, "mean", True).backward()

            t_torch, t_ours = _timed_ms(torch_step, flush, n=5, warm=2), _timed_ms(our_step, flush, n=5, warm=2)
            res[f"torch_cuda_ctc_{name}"] = {"torch_ms": t_torch, "ours_ms": t_ours, "speedup": t_torch / t_ours,
                                              "torch_cells_per_s": cells / (t_torch * 1e-3),
                                              "ours_cells_per_s": cells / (t_ours * 1e-3)}
        except Exception as e:
            res[f"torch_cuda_ctc_{name}"] = {"error": repr(e)}
    # ---- f-1: log_softmax + ctc_loss (+ both backwards) against the single logits entry point, C5 shape
    try:
        lp_d, tg, il, tl, cells = batches["c5"]
        logits = lp_d * 1.5 + 2.0
        tg_d, il_d, tl_d = tg.to(dev, torch.int32), il.to(dev, torch.int32), tl.to(dev, torch.int32)

        def unfused():
            x = logits.detach().requires_grad_(True)
            ssak_b200.ctc_loss(torch.log_softmax(x, -1), tg_d, il_d, tl_d, 0, "mean", True).backward()

        def fused():
            x = logits.detach().requires_grad_(True)
            ssak_b200.ctc_loss_from_logits(x, tg_d, il_d, tl_d, 0, "mean", True).backward()

        tms = {}
        for nm, fn in (("log_softmax_then_ctc_ms", unfused), ("from_logits_ms", fused)):
            tms[nm] = _timed_ms(fn, flush, n=5, warm=2)
        tms["cells_per_s_from_logits"] = cells / (tms["from_logits_ms"] * 1e-3)
        sumTV = float(il.sum()) * 1024
        tms["hbm_frac_16B"] = 16.0 * sumTV / (tms["from_logits_ms"] * 1e-3) / 1e9 / hbm
        res["loss_c5_logits"] = tms
        del logits
    except Exception as e:
        res["loss_c5_logits"] = {"error": repr(e)}
    batches.clear()
    torch.cuda.empty_cache()
    # ---- forced alignment (C5, C2-shaped, C3) and greedy
    ctx = C.c_void_p()
    assert lib.ssak_context_create(dev.index or 0, C.byref(ctx)) == 0
    for name, (B, T, V, Lmin, Lmax, Tmin) in {"align_c5": (512, 750, 1024, 100, 200, 600),
                                               "align_c2shape": (64, 1500, 50, 200, 400, 1200),
                                               "align_1k": (1024, 1500, 50, 200, 400, 1200),
                                               "align_c3": (16, 30000, 50, 7600, 8000, 30000)}.items():
        try:
            em, toks, el, tl = align_batch(B, T, V, Lmin, Lmax, 5, Tmin=Tmin)
            em_d, toks_d, el_d, tl_d = em.to(dev), toks.to(dev), el.to(dev), tl.to(dev)
            keep = {}

            def run_align():
                keep["r"] = ssak_b200.forced_align(em_d, toks_d, el_d, tl_d)

            t = _timed_ms(run_align, flush, n=5, warm=2) * 1e-3
            r = keep["r"]
            cells = int((el.long() * (tl.long() + 1)).sum())
            alg_bytes = (4.0 * float(el.sum()) * V + 4.0 * float((el.long() * (tl.long() + 1)).sum()) / 8
                         + 16.0 * float(tl.sum()))
            ach = alg_bytes / t / 1e9
            res[name] = {"cells_per_s": cells / t, "ms": t * 1e3, "aligned": int((r.status == 0).sum()), "B": B,
                         "roofline": {"bound": "hbm", "kernel": "align_wave_kernel + align_backtrace_kernel (one C call)",
                                      "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": t * 1e3, "achieved": ach,
                                      "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic_of(name),
                                      "peak_source": hbm_src}}
            # end to end: host emissions in, spans / scores out (ssak_forced_align_host), pinned host buffers
            em_h = em.pin_memory()
            tk_h = toks.contiguous()
            Lm = tk_h.shape[1]
            st = torch.empty((B, Lm), dtype=torch.int32).pin_memory()
            en = torch.empty((B, Lm), dtype=torch.int32).pin_memory()
            sc = torch.empty((B, Lm), dtype=torch.float64).pin_memory()
            ts, status = torch.empty(B, dtype=torch.int32), torch.empty(B, dtype=torch.int32)

            def host_align():
                rc = lib.ssak_forced_align_host(ctx, em_h.data_ptr(), B, T, V, tk_h.data_ptr(), Lm, el.data_ptr(),
                                                tl.data_ptr(), 0, 0, None, st.data_ptr(), en.data_ptr(), sc.data_ptr(),
                                                ts.data_ptr(), status.data_ptr())
                assert rc == 0, rc

            host_align()
            t0 = time.perf_counter()
            for _ in range(3):
                host_align()
            te = (time.perf_counter() - t0) / 3
            assert torch.equal(st[:, :Lm], r.starts.cpu()) and int((status == 0).sum()) == B
            res[name]["e2e"] = {"cells_per_s": cells / te, "ms": te * 1e3, "h2d_bytes": em.numel() * 4 + tk_h.numel() * 4 + 8 * B,
                                "d2h_bytes": B * Lm * 16 + 8 * B, "path": "ssak_forced_align_host (C ABI)"}
            if name == "align_c2shape":   # the CPU port (oracle C, one core) on a bounded sample of the same batch
                from oracle import oracle as O
                t0, n_cpu, c_cpu = time.perf_counter(), 0, 0
                while n_cpu < B and time.perf_counter() - t0 < 5.0:
                    Tb, Lb = int(el[n_cpu]), int(tl[n_cpu])
                    O.align(em[n_cpu, :Tb].numpy(), toks[n_cpu, :Lb].tolist(), 0, False)
                    c_cpu += Tb * (Lb + 1)
                    n_cpu += 1
                res[name]["cpu_port"] = {"cells_per_s": c_cpu / (time.perf_counter() - t0), "cores": 1,
                                         "sample": f"{n_cpu} utterances of this batch, oracle/ssak_oracle.c"}
            if name == "align_c5":   # greedy decode of the same emissions: frames/s and fraction of the HBM roofline
                t = _timed_ms(lambda: ssak_b200.greedy_ids(em_d, el_d, 0), flush, n=5, warm=2) * 1e-3
                gb = 4.0 * B * T * V + 8.0 * B * T
                res["greedy_c5"] = {"frames_per_s": B * T / t, "ms": t * 1e3,
                                    "roofline": {"bound": "hbm", "kernel": "greedy_argmax_kernel + greedy_collapse_kernel",
                                                 "algorithmic_bytes_per_launch": gb, "launch_ms": t * 1e3,
                                                 "achieved": gb / t / 1e9, "peak": hbm, "unit": "GB/s",
                                                 "frac": gb / t / 1e9 / hbm, "traffic": None}}
            del em_d, em_h
            torch.cuda.empty_cache()
        except Exception as e:
            res[name] = {"error": repr(e)}
    lib.ssak_context_destroy(ctx)
    return res


def main():
    if os.environ.get("BENCH_WATCHDOG"):   # debugging aid: dump every thread's stack and exit after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1k", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: one batch of the workload per GPU; strong: ONE global batch split by lpt_partition")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (rank 0, N = 1)")
    ap.add_argument("--extra", action="store_true", help="(default now; kept for compatibility)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
