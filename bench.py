#!/usr/bin/env python
"""bench.py -- lattice-cells/s of the CTC loss forward+backward hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path, rank 0)

Workload (BASELINE.json configs[1], "C2"): synthetic CTC loss fwd+bwd, B=64 utterances per GPU,
T=1500 frames, V=50, L in [200,400], T_b in [1200,1500], fp32, planted-alignment emissions
(SURVEY.md section 8d).  A step = one forward + one backward of the loss over the batch.
cells = sum_b T_b*(2L_b+1), counted once per fwd+bwd.  Prints ONE JSON line (rank 0).
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
    "c5": (512, 750, 1024, 100, 200, 600),
}
METRIC = "ctc_lattice_cells_per_sec"
UNIT = "cells/s"


def make_batch(name, seed):
    from ssak_b200.synth import ctc_batch
    B, T, V, Lmin, Lmax, Tmin = WORKLOADS[name]
    lp, tg, il, tl = ctc_batch(B, T, V, Lmin, Lmax, seed, Tmin=Tmin, planted=True)
    cells = int((il * (2 * tl + 1)).sum())
    return lp, tg, il, tl, cells


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


def run_reference(args, rank):
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    torch.set_num_threads(ncores)
    lp, tg, il, tl, cells = make_batch(args.workload, 1234 + 2)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_time(lp, tg, il, tl, 1)
    times = [cpu_reference_time(lp, tg, il, tl, 1) for _ in range(args.steps)]
    t = sum(times) / len(times)
    B, T, V, _, Lmax, _ = WORKLOADS[args.workload]
    val = cells / t
    sample = f"full {args.workload} batch (B={B}) per step, torch {torch.__version__} CPU F.ctc_loss fwd+bwd"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: CTC loss fwd+bwd B={B} T={T} V={V} L<={Lmax} (rank-0 host only)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def time_kernels(lib, dev, lp_d, tg32, off, il32, tl32, Lmax, iters, flush):
    """CUDA-event time of the forward launch pair and of the backward launch, separately."""
    T, B, V = lp_d.shape
    ws_bytes = lib.ssak_ctc_loss_workspace_bytes(T, B, Lmax, 1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    nll = torch.empty(B, dtype=torch.float32, device=dev)
    go = torch.full((B,), 1.0 / B, dtype=torch.float32, device=dev)
    grad = torch.empty_like(lp_d)
    s = torch.cuda.current_stream().cuda_stream
    tf, tb = [], []
    for i in range(iters + 2):
        flush.add_(1)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        rc = lib.ssak_ctc_loss_forward(lp_d.data_ptr(), T, B, V, lp_d.stride(0), lp_d.stride(1), tg32.data_ptr(),
                                       off.data_ptr(), il32.data_ptr(), tl32.data_ptr(), Lmax, 0, 1, nll.data_ptr(),
                                       ws.data_ptr(), ws_bytes, s)
        assert rc == 0
        e1.record()
        rc = lib.ssak_ctc_loss_backward(go.data_ptr(), lp_d.data_ptr(), T, B, V, lp_d.stride(0), lp_d.stride(1),
                                        tg32.data_ptr(), off.data_ptr(), il32.data_ptr(), tl32.data_ptr(), Lmax, 0, 1,
                                        nll.data_ptr(), grad.data_ptr(), grad.stride(0), grad.stride(1),
                                        ws.data_ptr(), ws_bytes, s)
        assert rc == 0
        e2.record()
        torch.cuda.synchronize()
        if i >= 2:
            tf.append(e0.elapsed_time(e1))
            tb.append(e1.elapsed_time(e2))
    return statistics.mean(tf) * 1e-3, statistics.mean(tb) * 1e-3


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import ssak_b200
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = ssak_b200.lib()
    name = args.workload
    B, T, V, Lmin, Lmax, Tmin = WORKLOADS[name]
    lp, tg, il, tl, cells = make_batch(name, 1234 + 2 + 1000 * rank)   # weak scaling: own batch per rank
    lp_pin, grad_pin = lp.pin_memory(), torch.empty_like(lp).pin_memory()
    lp_d = lp_pin.to(dev, non_blocking=True)
    tg_d, il_d, tl_d = tg.to(dev, torch.int32), il.to(dev, torch.int32), tl.to(dev, torch.int32)
    flush = torch.zeros(96 * 1024 * 1024, dtype=torch.float32, device=dev)      # 384 MB > 126 MB L2

    def step():
        x = lp_d.detach().requires_grad_(True)
        if world > 1:   # utterance-sharded batch: local lattices + ONE all-reduce of 2 scalars (NCCL)
            from ssak_b200.shard import sharded_ctc_loss
            loss = sharded_ctc_loss(x, tg_d, il_d, tl_d, blank=0, reduction="mean", zero_infinity=True,
                                    global_batch=B * world)
        else:
            loss = ssak_b200.ctc_loss(x, tg_d, il_d, tl_d, blank=0, reduction="mean", zero_infinity=True)
        loss.backward()
        return loss, x.grad

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
    # the timed region lasts a few ms, shorter than nvidia-smi's sampling period: keep the same work running
    # for ~0.3 s more so that the clock / throttle samples are taken under this load
    # (every rank runs the same fixed number of steps: the sharded step contains a collective)
    for _ in range(30):
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    # ---- end to end through the host-buffer C ABI: pinned host log-probs in, nll + gradient out
    ctx = C.c_void_p()
    assert lib.ssak_context_create(local_rank, C.byref(ctx)) == 0
    tg32 = tg.to(torch.int32).contiguous()
    il32, tl32 = il.to(torch.int32), tl.to(torch.int32)
    nll_h = torch.empty(B, dtype=torch.float32).pin_memory()

    def e2e_step():
        rc = lib.ssak_ctc_loss_host(ctx, lp_pin.data_ptr(), T, B, V, tg32.data_ptr(), tg32.shape[1],
                                    il32.data_ptr(), tl32.data_ptr(), 0, 1, None, nll_h.data_ptr(),
                                    grad_pin.data_ptr())
        assert rc == 0, rc

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    if rank == 0:
        sampler.stop()
    # parity gate on the timed configuration: the host-ABI result equals the torch-facing one
    _, g = step()
    torch.cuda.synchronize()
    gscale = (1.0 / (B * world * tl.clamp_min(1).float())).view(1, B, 1)
    assert (g.cpu() - grad_pin * gscale).abs().max().item() < 1e-6

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
        off = (torch.arange(B, device=dev, dtype=torch.int64) * tg.shape[1])
        t_fwd, t_bwd = time_kernels(lib, dev, lp_d, tg32.to(dev), off, il32.to(dev), tl32.to(dev), int(tl.max()),
                                    max(args.steps, 5), flush)
        sumTV = float(il.sum()) * V
        bwd_bytes = 8.0 * sumTV            # read every emission row once + write the gradient row
        achieved = bwd_bytes / t_bwd / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(name)
        cpu_s = cpu_reference_time(lp, tg, il, tl, 3)
        ncores = os.cpu_count() or 1
        h2d = lp.numel() * 4 + tg32.numel() * 4 + 2 * B * 4 + B * 8 + B * 4
        d2h = lp.numel() * 4 + B * 4
        out = {
            "metric": METRIC, "value": total * args.steps / t_dev, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{name}: CTC loss fwd+bwd, B={B} per GPU, T={T}, V={V}, L in [{Lmin},{Lmax}], "
                                   f"T_b in [{Tmin},{T}], planted-alignment emissions, reduction=mean, zero_infinity",
                       "cells_per_step_per_gpu": cells, "l2": "flushed (384 MB write) before every timed step",
                       "inputs": "log_probs fp32 [T,B,V] contiguous, int32 padded targets / lengths, resident in HBM",
                       "timing": "per-step CUDA events on the launching stream, summed; max over ranks"},
            "e2e": {"value": total * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e / args.steps * 1e3,
                    "path": "ssak_ctc_loss_host (C ABI): pinned host log-probs in, nll + full gradient out"},
            "gpu_launches": 4 * args.steps,   # lattice fwd, join, reduce, lattice bwd
            "roofline": {"bound": "hbm", "kernel": "ctc_lattice_kernel<K,true> (backward: recursion + gradient)",
                         "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "peak_source": hbm_src,
                         "algorithmic_bytes_per_launch": bwd_bytes, "launch_ms": t_bwd * 1e3,
                         "forward_launch_ms": t_fwd * 1e3,
                         "forward_achieved_gbs": 4.0 * sumTV / t_fwd / 1e9,
                         "note": "V=50: 0.75 B/cell of compulsory traffic, the kernel is bound by the serial "
                                 "recursion and MUFU (DESIGN.md), not by HBM"},
            "cpu_baseline": {"value": cells / cpu_s, "unit": UNIT, "cores": ncores, "kind": "reference",
                             "sample": f"full {name} batch (B={B}) fwd+bwd, best of 3, torch {torch.__version__} "
                                       f"CPU F.ctc_loss with {ncores} threads"},
            "clocks": sampler.summary(),
        }
        if args.extra:
            out["extra"] = extra_numbers(lib, dev, flush)
    lib.ssak_context_destroy(ctx)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


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


def extra_numbers(lib, dev, flush):
    """Other configurations of BASELINE.json, device-resident, kernels only (informational)."""
    import ssak_b200
    from ssak_b200.synth import align_batch
    res = {}
    for name in ("1k", "c5"):
        try:
            B, T, V, Lmin, Lmax, Tmin = WORKLOADS[name]
            lp, tg, il, tl, cells = make_batch(name, 99)
            lp_d = lp.to(dev)
            off = torch.arange(B, device=dev, dtype=torch.int64) * tg.shape[1]
            tf, tb = time_kernels(lib, dev, lp_d, tg.to(torch.int32).to(dev), off, il.to(torch.int32).to(dev),
                                  tl.to(torch.int32).to(dev), int(tl.max()), 5, flush)
            sumTV = float(il.sum()) * V
            res[f"loss_{name}"] = {"cells_per_s": cells / (tf + tb), "fwd_ms": tf * 1e3, "bwd_ms": tb * 1e3,
                                   "hbm_frac_canonical_12B": 12.0 * sumTV / (tf + tb) / 1e9 / peaks()[0]}
            del lp_d
        except Exception as e:  # keep the headline line even if an extra shape fails
            res[f"loss_{name}"] = {"error": repr(e)}
    # C4 (BASELINE config 4): B=256 utterances with T_b ~ U[300,1500], L_b = 0.27 T_b, V=50 -- one launch over the
    # whole ragged batch against 4 length buckets (ssak_b200.shard.length_buckets; the batches a bucketed sampler
    # would hand over), fwd+bwd kernels, one GPU's share of the 2/4/8-GPU configuration
    try:
        from ssak_b200.shard import lattice_cost, length_buckets
        from ssak_b200.synth import planted_emissions
        g = torch.Generator().manual_seed(1234 + 4)
        B4, T4, V4 = 256, 1500, 50
        il = torch.randint(300, T4 + 1, (B4,), generator=g)
        tl = (0.27 * il.float()).round().long().clamp_min(1)
        tg = torch.randint(1, V4, (B4, int(tl.max())), generator=g)
        lp = torch.empty(T4, B4, V4)
        for b in range(B4):
            e = torch.randn(T4, V4, generator=g)
            e[: int(il[b])] = planted_emissions(int(il[b]), V4, tg[b, : int(tl[b])], g, 0, normalize=False)
            lp[:, b] = e.log_softmax(-1)
        cells = int((il * (2 * tl + 1)).sum())

        def timed(idx):
            idx = torch.as_tensor(idx)
            Tm = int(il[idx].max())
            sub = lp[:Tm, idx].contiguous().to(dev)
            off = torch.arange(len(idx), device=dev, dtype=torch.int64) * tg.shape[1]
            tf, tb = time_kernels(lib, dev, sub, tg[idx].to(torch.int32).to(dev), off, il[idx].to(torch.int32).to(dev),
                                  tl[idx].to(torch.int32).to(dev), int(tl[idx].max()), 5, flush)
            return tf + tb

        t_all = timed(list(range(B4)))
        t_bkt = sum(timed(bk) for bk in length_buckets(lattice_cost(il.tolist(), tl.tolist()), 4))
        res["loss_c4"] = {"cells_per_s_one_launch": cells / t_all, "ms_one_launch": t_all * 1e3,
                          "cells_per_s_4_length_buckets": cells / t_bkt, "ms_4_length_buckets": t_bkt * 1e3}
        del lp
    except Exception as e:
        res["loss_c4"] = {"error": repr(e)}
    # on-box GPU comparator: torch's own CUDA ctc_loss (native kernel, cuDNN off as HF does) on the
    # same C2 / 1k / C5 tensors, forward + backward through autograd, same event timing and L2 flush
    import torch.nn.functional as F
    for name in ("c2", "1k", "c5"):
        try:
            B, T, V, Lmin, Lmax, Tmin = WORKLOADS[name]
            lp, tg, il, tl, cells = make_batch(name, 99)
            lp_d, tg_d, il_d, tl_d = lp.to(dev), tg.to(dev), il.to(dev), tl.to(dev)

            def torch_step():
                x = lp_d.detach().requires_grad_(True)
                with torch.backends.cudnn.flags(enabled=False):
                    F.ctc_loss(x, tg_d, il_d, tl_d, 0, "mean", True).backward()

            def our_step():
                x = lp_d.detach().requires_grad_(True)
                ssak_b200.ctc_loss(x, tg_d, il_d, tl_d, 0, "mean", True).backward()

            t_torch, t_ours = _timed_ms(torch_step, flush), _timed_ms(our_step, flush)
            res[f"torch_cuda_ctc_{name}"] = {"torch_ms": t_torch, "ours_ms": t_ours, "speedup": t_torch / t_ours,
                                              "torch_cells_per_s": cells / (t_torch * 1e-3)}
            del lp_d
        except Exception as e:
            res[f"torch_cuda_ctc_{name}"] = {"error": repr(e)}
    # f-1: log_softmax + ctc_loss (+ both backwards) against the single logits entry point, C5 shape, torch-facing API
    try:
        B, T, V, Lmin, Lmax, Tmin = WORKLOADS["c5"]
        lp, tg, il, tl, cells = make_batch("c5", 99)
        logits = (lp * 1.5 + 2.0).to(dev)
        tg_d, il_d, tl_d = tg.to(dev, torch.int32), il.to(dev, torch.int32), tl.to(dev, torch.int32)
        del lp

        def unfused():
            x = logits.detach().requires_grad_(True)
            ssak_b200.ctc_loss(torch.log_softmax(x, -1), tg_d, il_d, tl_d, 0, "mean", True).backward()

        def fused():
            x = logits.detach().requires_grad_(True)
            ssak_b200.ctc_loss_from_logits(x, tg_d, il_d, tl_d, 0, "mean", True).backward()

        tms = {}
        for nm, fn in (("log_softmax_then_ctc_ms", unfused), ("from_logits_ms", fused)):
            tms[nm] = _timed_ms(fn, flush)
        tms["cells_per_s_from_logits"] = cells / (tms["from_logits_ms"] * 1e-3)
        res["loss_c5_logits"] = tms
        del logits
    except Exception as e:
        res["loss_c5_logits"] = {"error": repr(e)}
    for name, (B, T, V, Lmin, Lmax, Tmin) in {"align_c5": (512, 750, 1024, 100, 200, 600),
                                               "align_c2shape": (64, 1500, 50, 200, 400, 1200),
                                               "align_c3": (16, 30000, 50, 7600, 8000, 30000)}.items():
        try:
            em, toks, el, tl = align_batch(B, T, V, Lmin, Lmax, 5, Tmin=Tmin)
            em_d, toks_d, el_d, tl_d = em.to(dev), toks.to(dev), el.to(dev), tl.to(dev)
            keep = {}

            def run_align():
                keep["r"] = ssak_b200.forced_align(em_d, toks_d, el_d, tl_d)

            t = _timed_ms(run_align, flush) * 1e-3
            r = keep["r"]
            cells = int((el.long() * (tl.long() + 1)).sum())
            alg_bytes = 4.0 * float(el.sum()) * V + 4.0 * float((el.long() * (tl.long() + 1)).sum()) / 8 * 2
            res[name] = {"cells_per_s": cells / t, "ms": t * 1e3, "aligned": int((r.status == 0).sum()),
                         "hbm_frac_algorithmic": alg_bytes / t / 1e9 / peaks()[0]}
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
                t = _timed_ms(lambda: ssak_b200.greedy_ids(em_d, el_d, 0), flush) * 1e-3
                res["greedy_c5"] = {"frames_per_s": B * T / t, "ms": t * 1e3,
                                    "hbm_frac": (4.0 * B * T * V + 8.0 * B * T) / t / 1e9 / peaks()[0]}
            del em_d
        except Exception as e:
            res[name] = {"error": repr(e)}
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
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--extra", action="store_true", help="also time the other BASELINE configs (rank 0)")
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
