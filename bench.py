#!/usr/bin/env python
"""Benchmark of the HiFi-GAN decoder hot path (BASELINE.json metric: decoder audio-seconds/second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic latents:
  train_base_b16 (default, BASELINE.json configs[1]): configs/base.json decoder, batch 16 x 32-frame segments,
      bf16 tensor-core mode: weight-norm fold + forward + backward (all 233 parameter gradients, dz, dg).
  train_48k_b32 (configs[2]): 48k_base.json (inter_channels 128), batch 32 per GPU, same step.
  infer_10s (configs[3], scaled by --infer-batch): forward only on 938-frame (10 s) latents.
N > 1 (torchrun): one rank per GPU, identical weights, different latents per rank, gradient all-reduce (NCCL)
overlapped with backward inside the step; value = audio-seconds of ALL ranks / max-over-ranks step time.

One JSON line is printed by rank 0 (see the contract in the task description): `value` is device-resident
throughput, `e2e` the same step fed from pinned host memory with the waveform read back, `roofline` the
dominant kernel class (tcgen05 implicit-GEMM convolutions) against the measured bf16 peak, `cpu_baseline`
the CPU oracle timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    #                cfg name      B   T    train
    "train_base_b16": ("BASE_CFG", 16, 32, True),
    "train_48k_b32": ("BASE48K_CFG", 32, 32, True),
    "infer_10s": ("BASE_CFG", 8, 938, False),
    "fwd_base_b1": ("BASE_CFG", 1, 32, False),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train_base_b16", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--infer-batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--profile-classes", action="store_true", help="print the per-kernel-class table to stderr")
    ap.add_argument("--dump-launches", default=None, help="write the per-launch CSV of the profiled steps here")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm": d["hbm_gbs"], "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (PyTorch CPU restatement of the reference decoder) on host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_step_time(cfg, B, T, train, steps, warmup, seed=1234):
    import torch
    from oracle import hifigan_oracle as O
    sd = O.seeded_state_dict(cfg, seed)
    model = O.build(cfg, sd)
    torch.manual_seed(seed)
    x = torch.randn(B, cfg["initial_channel"], T)
    g = torch.randn(B, cfg["gin_channels"], 1)
    hop = 1
    for u in cfg["upsample_rates"]:
        hop *= u
    dy = torch.randn(B, 1, T * hop)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if train:
            model.zero_grad(set_to_none=True)
            xx = x.clone().requires_grad_(True)
            gg = g.clone().requires_grad_(True)
            y = model(xx, gg)
            y.backward(dy)
        else:
            with torch.no_grad():
                model(x, g)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times), torch.get_num_threads()


def run_reference(args):
    import torch
    from oracle import hifigan_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_name, B, T, train = WORKLOADS[args.workload]
    if args.workload == "infer_10s" and args.infer_batch:
        B = args.infer_batch
    cfg = getattr(O, cfg_name)
    # bounded sample of the workload: the CPU step is slow, keep the whole run within a few minutes
    steps = max(1, min(args.steps, 3))
    warmup = max(1, min(args.warmup, 1))
    sB = B if (train and T * B <= 512) else max(1, min(B, 2))
    t, cores = cpu_step_time(cfg, sB, T, train, steps, warmup)
    value = O.audio_seconds(sB, T) / t
    sample = f"{'fwd+bwd' if train else 'fwd'} of {cfg_name} B={sB} T={T} fp32, {steps} timed steps after {warmup} warm-up"
    line = {"impl": "reference", "metric": "decoder_audio_seconds_per_second", "value": value, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "cfg": cfg_name, "batch_per_step": sB, "frames": T,
                       "note": "CPU oracle (PyTorch restatement of the reference decoder; the reference ships no "
                               "Generator class) on host cores; rank 0 only"},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from oracle import hifigan_oracle as O
    from vcvits_b200 import Generator, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg_name, B, T, train = WORKLOADS[args.workload]
    if args.workload == "infer_10s" and args.infer_batch:
        B = args.infer_batch
    cfg = getattr(O, cfg_name)
    lib = _lib.load()

    torch.manual_seed(1234)  # identical weights on every rank (configs/base.json:12)
    model = Generator(**cfg, mode=args.mode).to(dev)
    if world > 1 and train:
        model.set_gradient_sync(dist.group.WORLD)
    hop = model.hop
    gen = torch.Generator(device="cpu").manual_seed(1000 + rank)  # different data per rank
    x_host = torch.randn(B, cfg["initial_channel"], T, generator=gen).pin_memory()
    g_host = torch.randn(B, cfg["gin_channels"], 1, generator=gen).pin_memory()
    dy_host = torch.randn(B, 1, T * hop, generator=gen).pin_memory()
    y_host = torch.empty(B, 1, T * hop).pin_memory()
    x_dev, g_dev, dy_dev = x_host.to(dev), g_host.to(dev), dy_host.to(dev)
    flush_buf = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    params = list(model.parameters())

    def step_device():
        if train:
            for p in params:
                p.grad = None
            model._fold_key = None  # parameters change every optimizer step: the fold is part of the step
            xx = x_dev.detach().requires_grad_(True)
            gg = g_dev.detach().requires_grad_(True)
            y = model(xx, gg)
            y.backward(dy_dev)
            return y
        with torch.no_grad():
            return model(x_dev, g_dev)

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        gd = g_host.to(dev, non_blocking=True)
        if train:
            dyd = dy_host.to(dev, non_blocking=True)
            for p in params:
                p.grad = None
            model._fold_key = None
            y = model(xd.requires_grad_(True), gd.requires_grad_(True))
            y.backward(dyd)
        else:
            with torch.no_grad():
                y = model(xd, gd)
        y_host.copy_(y.detach(), non_blocking=True)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        host = 0.0
        for _ in range(steps):
            if flush_buf is not None:
                flush_buf.fill_(1)  # evict L2 (256 MB write > 126 MB L2); outside the timed events
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            h0 = time.perf_counter()
            fn()
            host += time.perf_counter() - h0
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        timed.host_ms = host / steps * 1e3
        return float(t.item()) / steps

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.vcd_launch_count(1)
    ms = timed(step_device, args.steps, args.warmup)
    host_ms = timed.host_ms
    launches_total = lib.vcd_launch_count(1)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = launches_total / (args.steps + args.warmup)

    e2e = None
    if not args.no_e2e:
        ms_e2e = timed(step_e2e, args.steps, max(3, args.warmup // 2))
        h2d = x_host.numel() * 4 + g_host.numel() * 4 + (dy_host.numel() * 4 if train else 0)
        e2e = {"value": world * O.audio_seconds(B, T) / (ms_e2e * 1e-3), "unit": "audio-s/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": y_host.numel() * 4, "ms_per_step": ms_e2e}

    # ---- roofline of the dominant kernel class, measured live with CUDA events around every launch ----
    peaks = load_peaks()
    roofline, classes = None, None
    if world > 1:
        dist.barrier()
    if rank == 0:
        model.set_gradient_sync(None)  # rank-0-only profiling passes must not enter a collective
        lib.vcd_profile_enable(1)
        for _ in range(2):
            step_device()
        n = lib.vcd_profile_num_classes()
        arr_ms, arr_l = (C.c_double * n)(), (C.c_uint64 * n)()
        arr_f, arr_b = (C.c_double * n)(), (C.c_double * n)()
        lib.vcd_profile_read(1, arr_ms, arr_l, arr_f, arr_b)
        psteps = 3
        for _ in range(psteps):
            if flush_buf is not None:
                flush_buf.fill_(1)
            step_device()
        if args.dump_launches:
            lib.vcd_profile_dump(args.dump_launches.encode())
        lib.vcd_profile_read(1, arr_ms, arr_l, arr_f, arr_b)
        lib.vcd_profile_enable(0)
        classes = []
        for c in range(n):
            if arr_l[c]:
                classes.append({"class": lib.vcd_profile_class_name(c).decode(), "ms_per_step": arr_ms[c] / psteps,
                                "launches_per_step": arr_l[c] / psteps, "gflop_per_step": arr_f[c] / psteps / 1e9,
                                "tflops": arr_f[c] / (arr_ms[c] * 1e-3) / 1e12 if arr_ms[c] > 0 else 0.0})
        dom = max(classes, key=lambda d: d["ms_per_step"]) if classes else None
        tcs = [c for c in classes if c["class"].startswith("tc_")]
        if tcs:
            fl = sum(c["gflop_per_step"] for c in tcs) * 1e9
            tm = sum(c["ms_per_step"] for c in tcs) * 1e-3
            ln = sum(c["launches_per_step"] for c in tcs)
            achieved = fl / tm / 1e12
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
            if args.workload == "train_base_b16" and os.path.exists(tpath):
                with open(tpath) as tf:
                    traffic = json.load(tf).get("dram_bytes_per_launch")
            roofline = {"bound": "tensor", "kernel": "tc::conv_kernel / tc::wgrad_kernel (tcgen05 implicit GEMM)",
                        "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                        "frac": achieved / peaks["bf16_sustained"], "traffic": traffic,
                        "peak_source": f"{peaks['source']} sustained bf16 (kernel timed inside a long step)",
                        "gflop_per_launch": fl / ln / 1e9, "avg_launch_us": tm / ln * 1e6,
                        "share_of_step": tm / (sum(c["ms_per_step"] for c in classes) * 1e-3),
                        "dominant_class": dom["class"] if dom else None}
        if args.profile_classes:
            for c in classes:
                print(json.dumps(c), file=sys.stderr)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sB = B if (train and B * T <= 512) else max(1, min(B, 2))
        t_cpu, cores = cpu_step_time(cfg, sB, T, train, 2, 1)
        cpu_baseline = {"value": O.audio_seconds(sB, T) / t_cpu, "unit": "audio-s/s", "cores": cores, "kind": "port",
                        "sample": f"{'fwd+bwd' if train else 'fwd'} of {cfg_name} B={sB} T={T} fp32 CPU oracle, "
                                  f"2 timed steps after 1 warm-up ({t_cpu:.2f} s/step)"}

    if rank == 0:
        audio_s = world * O.audio_seconds(B, T)
        flops = O.forward_flops(cfg, B, T) * (3 if train else 1)
        value = audio_s / (ms * 1e-3)
        line = {"metric": "decoder_audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.mode, "data": "synthetic",
                "config": {"workload": args.workload, "cfg": cfg_name, "batch_per_gpu": B, "frames": T,
                           "step": "weight-norm fold + fwd + bwd (233 param grads, dz, dg)" + (" + NCCL grad all-reduce" if world > 1 else "") if train else "fwd",
                           "l2": "256 MB L2 flush between timed steps" if flush_buf is not None else "no flush",
                           "random_init_weights": True},
                "tflops_algorithmic": flops / (ms * 1e-3) / 1e12,
                "frac_of_bf16_peak_sustained": flops / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
                "host_enqueue_ms_per_step": host_ms,
                "gpu_launches": round(launches_per_step * args.steps), "gpu_launches_per_step": launches_per_step,
                "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "kernel_classes": classes}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # Only the JSON line may reach stdout: libraries (e.g. NCCL's version banner) write to fd 1 directly, so fd 1 is
    # pointed at stderr for the whole run and the result is written to the saved descriptor.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved, "w")
    import builtins
    _print = builtins.print

    def emit(*a, **k):
        if k.get("file") is None and a and isinstance(a[0], str) and a[0].startswith("{"):
            _print(*a, file=real_stdout, flush=True)
        else:
            _print(*a, **k)

    builtins.print = emit
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    finally:
        builtins.print = _print
        real_stdout.flush()


if __name__ == "__main__":
    main()
