#!/usr/bin/env python
"""Benchmark of the HiFi-GAN decoder hot path (BASELINE.json metric: decoder audio-seconds/second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic latents:
  train_base_b16 (default, BASELINE.json configs[1]): configs/base.json decoder, batch 16 x 32-frame segments,
      bf16 tensor-core mode: weight-norm fold + forward + backward (all 233 parameter gradients, dz, dg).
  train_48k_b32 (configs[2]): 48k_base.json (inter_channels 128), batch 32 per GPU, same step.
  infer_10s (configs[3]): forward only on 938-frame (10 s) latents, batch 64 per GPU (--infer-batch overrides).
  fwd_base_b1 (configs[0]'s shape on the GPU): forward, batch 1 x 32 frames.
  full_synth (configs[4]): decoder share of a full synthesizer forward -- see tools/full_synth.py.
N > 1 (torchrun): one rank per GPU, identical weights, different latents per rank, gradient all-reduce (NCCL)
overlapped with backward inside the step; value = audio-seconds of ALL ranks / max-over-ranks step time.

One JSON line is printed by rank 0 (see the contract in the task description): `value` is device-resident
throughput, `e2e` the same step fed from pinned host memory with the waveform read back, `roofline` the tcgen05
implicit-GEMM kernels against the measured bf16 peak (`roofline_classes` lists every kernel class, the HBM-bound
ones against the measured copy bandwidth), `cpu_baseline` the CPU oracle timed on this box's host cores, and
`configs` short sub-records of the other BASELINE.json GPU configurations (configs[2], configs[3]) at the same N.
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
    "infer_10s": ("BASE_CFG", 64, 938, False),
    "fwd_base_b1": ("BASE_CFG", 1, 32, False),
}
METRIC = "decoder_audio_seconds_per_second"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train_base_b16", choices=sorted(WORKLOADS) + ["full_synth"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--infer-batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[2] / configs[3] sub-records")
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


def workload_shape(args, name):
    cfg_name, B, T, train = WORKLOADS[name]
    if name == "infer_10s" and args.infer_batch:
        B = args.infer_batch
    return cfg_name, B, T, train


def make_config(args, name, world):
    """The `config` object -- identical in the b200 arm and the reference arm for the same command line."""
    cfg_name, B, T, train = workload_shape(args, name)
    step = "weight-norm fold + fwd + bwd (233 param grads, dz, dg)" if train else "fwd"
    if train and world > 1:
        step += " + gradient all-reduce"
    return {"workload": name, "cfg": cfg_name, "batch_per_gpu": B, "frames": T, "step": step,
            "l2": "no flush" if args.no_flush else "256 MB L2 flush between timed steps",
            "random_init_weights": True}


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
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must not inherit that."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_model(cfg, seed=1234):
    from oracle import hifigan_oracle as O
    return O.build(cfg, O.seeded_state_dict(cfg, seed))


def cpu_time_steps(model, cfg, B, T, train, steps, warmup, seed=1234):
    import torch
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
    return sum(times) / len(times)


def cpu_bounded_sample(cfg, B, T, train, steps, warmup, budget_s):
    """Batch size of the per-step sample so that `steps + warmup` CPU steps fit in about `budget_s` seconds: one probe
    step on a single utterance gives the cost per utterance (the decoder is independent per utterance)."""
    model = cpu_model(cfg)
    per_item = cpu_time_steps(model, cfg, 1, T, train, 1, 1)
    sB = int(max(1, min(B, (budget_s / max(1, steps + warmup)) / max(per_item, 1e-6))))
    return model, sB


def run_reference(args):
    from oracle import hifigan_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if args.workload == "full_synth":
        print(json.dumps({"impl": "reference", "unavailable": "full_synth is a decoder-share harness, not a throughput metric"}))
        return
    cores = use_all_host_threads()
    cfg_name, B, T, train = workload_shape(args, args.workload)
    cfg = getattr(O, cfg_name)
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    model, sB = cpu_bounded_sample(cfg, B, T, train, steps, warmup, budget_s=150.0)
    t = cpu_time_steps(model, cfg, sB, T, train, steps, warmup)
    value = O.audio_seconds(sB, T) / t
    sample = (f"{'fwd+bwd' if train else 'fwd'} of {cfg_name}, {sB} of the {B} utterances per step, T={T}, fp32, "
              f"{steps} timed steps after {warmup} warm-up, {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args, args.workload, world),
            "note": "CPU oracle (PyTorch restatement of the reference decoder; the reference ships no Generator class) on "
                    "the host cores; rank 0 only; throughput is per utterance, so a bounded sample of the batch is exact",
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from vcvits_b200 import _lib
        self.lib = _lib.load()
        self.flush_buf = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)
        self.peaks = load_peaks()

    def timed(self, fn, steps, warmup):
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs, host = [], 0.0
        for _ in range(steps):
            if self.flush_buf is not None:
                self.flush_buf.fill_(1)  # evict L2 (256 MB write > 126 MB L2); outside the timed events
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            h0 = time.perf_counter()
            fn()
            host += time.perf_counter() - h0
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([total_ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        self.host_ms = host / steps * 1e3
        return float(t.item()) / steps

    def setup(self, name):
        """Model + synthetic inputs of one workload; returns the two step functions."""
        torch, dist, args = self.torch, self.dist, self.args
        from oracle import hifigan_oracle as O
        from vcvits_b200 import Generator
        cfg_name, B, T, train = workload_shape(args, name)
        cfg = getattr(O, cfg_name)
        torch.manual_seed(1234)  # identical weights on every rank (configs/base.json:12)
        model = Generator(**cfg, mode=args.mode).to(self.dev)
        if self.world > 1 and train:
            model.set_gradient_sync(dist.group.WORLD)
        hop = model.hop
        gen = torch.Generator(device="cpu").manual_seed(1000 + self.rank)  # different data per rank
        x_host = torch.randn(B, cfg["initial_channel"], T, generator=gen).pin_memory()
        g_host = torch.randn(B, cfg["gin_channels"], 1, generator=gen).pin_memory()
        dy_host = torch.randn(B, 1, T * hop, generator=gen).pin_memory() if train else None
        y_host = torch.empty(B, 1, T * hop).pin_memory()
        x_dev, g_dev = x_host.to(self.dev), g_host.to(self.dev)
        dy_dev = dy_host.to(self.dev) if train else None
        params = list(model.parameters())
        dev = self.dev

        def step_device():
            if train:
                for p in params:
                    p.grad = None
                # parameters change every optimizer step: training-mode forwards always re-fold (part of the step)
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
                y = model(xd.requires_grad_(True), gd.requires_grad_(True))
                y.backward(dyd)
            else:
                with torch.no_grad():
                    y = model(xd, gd)
            y_host.copy_(y.detach(), non_blocking=True)

        h2d = x_host.numel() * 4 + g_host.numel() * 4 + (dy_host.numel() * 4 if train else 0)
        info = {"cfg": cfg, "cfg_name": cfg_name, "B": B, "T": T, "train": train, "model": model, "h2d": h2d,
                "d2h": y_host.numel() * 4}
        return step_device, step_e2e, info

    def profile_classes(self, step_device, model, psteps=3):
        """CUDA events around every launch (serial streams, no graphs): device ms / FLOPs / bytes per kernel class."""
        lib = self.lib
        model.set_gradient_sync(None)  # rank-0-only profiling passes must not enter a collective
        lib.vcd_profile_enable(1)
        for _ in range(2):
            step_device()
        n = lib.vcd_profile_num_classes()
        arr_ms, arr_l = (C.c_double * n)(), (C.c_uint64 * n)()
        arr_f, arr_b = (C.c_double * n)(), (C.c_double * n)()
        lib.vcd_profile_read(1, arr_ms, arr_l, arr_f, arr_b)
        for _ in range(psteps):
            if self.flush_buf is not None:
                self.flush_buf.fill_(1)
            step_device()
        if self.args.dump_launches:
            lib.vcd_profile_dump(self.args.dump_launches.encode())
        lib.vcd_profile_read(1, arr_ms, arr_l, arr_f, arr_b)
        lib.vcd_profile_enable(0)
        classes = []
        for c in range(n):
            if arr_l[c]:
                ms = arr_ms[c] / psteps
                classes.append({"class": lib.vcd_profile_class_name(c).decode(), "ms_per_step": ms,
                                "launches_per_step": arr_l[c] / psteps, "gflop_per_step": arr_f[c] / psteps / 1e9,
                                "gbyte_per_step": arr_b[c] / psteps / 1e9,
                                "tflops": arr_f[c] / (arr_ms[c] * 1e-3) / 1e12 if arr_ms[c] > 0 else 0.0,
                                "gbs": arr_b[c] / (arr_ms[c] * 1e-3) / 1e9 if arr_ms[c] > 0 else 0.0})
        return classes

    def rooflines(self, classes, workload):
        """Per kernel class: the roofline that bounds it follows from its arithmetic intensity (algorithmic FLOPs / algorithmic
        bytes, both accumulated per launch by the library) against the ridge of the MEASURED peaks (sustained bf16 / copy
        bandwidth, MEASURED_PEAKS.json).  The headline `roofline` is the class with the largest share of the step."""
        peaks = self.peaks
        ridge = peaks["bf16_sustained"] * 1e12 / (peaks["hbm"] * 1e9)    # FLOP per byte
        total_ms = sum(c["ms_per_step"] for c in classes)
        per_class = []
        for c in classes:
            if c["ms_per_step"] <= 0 or (c["gflop_per_step"] <= 0 and c["gbyte_per_step"] <= 0):
                continue
            ai = c["gflop_per_step"] / c["gbyte_per_step"] if c["gbyte_per_step"] > 0 else float("inf")
            rec = {"class": c["class"], "flop_per_byte": None if ai == float("inf") else ai,
                   "avg_launch_us": c["ms_per_step"] / c["launches_per_step"] * 1e3, "share_of_step": c["ms_per_step"] / total_ms,
                   "launches_per_step": c["launches_per_step"]}
            if c["gflop_per_step"] > 0 and ai >= ridge:
                rec.update({"bound": "tensor", "achieved": c["tflops"], "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                            "frac": c["tflops"] / peaks["bf16_sustained"]})
            else:
                rec.update({"bound": "hbm", "achieved": c["gbs"], "peak": peaks["hbm"], "unit": "GB/s", "frac": c["gbs"] / peaks["hbm"]})
                if c["gflop_per_step"] > 0:
                    rec["tflops"] = c["tflops"]
            per_class.append(rec)
        if not per_class:
            return None, per_class
        # DRAM bytes per launch cannot be measured without a profiler attached: taken from THIS round's committed
        # `ncu --set full` capture of the dominant class at this workload, or null when that capture is absent
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        dom = max(per_class, key=lambda d: d["share_of_step"])
        if os.path.exists(tpath):
            with open(tpath) as tf:
                td = json.load(tf)
            ent = td.get(workload, {}).get(dom["class"])
            if ent:
                traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
        kernel = {"tc_conv_c>=128(fwd+dgrad)": "tc::conv_kernel (tcgen05 implicit GEMM, streamed weights)",
                  "tc_conv_c<=64(fwd+dgrad,pair)": "tc::conv_kernel / tc::pair_kernel (tcgen05 implicit GEMM, resident weights)",
                  "tc_wgrad_c>=128": "tc::wgrad_kernel", "tc_wgrad_c<=64": "tc::wgrad_kernel"}.get(dom["class"], dom["class"])
        tcs = [c for c in classes if c["class"].startswith("tc_")]
        fl = sum(c["gflop_per_step"] for c in tcs) * 1e9
        tm = sum(c["ms_per_step"] for c in tcs) * 1e-3
        roofline = {"bound": dom["bound"], "kernel": kernel, "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"],
                    "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": f"{peaks['source']} ({'sustained bf16' if dom['bound'] == 'tensor' else 'copy bandwidth'}; kernel timed inside a long step)",
                    "flop_per_byte": dom["flop_per_byte"], "ridge_flop_per_byte": ridge,
                    "avg_launch_us": dom["avg_launch_us"], "share_of_step": dom["share_of_step"], "dominant_class": dom["class"],
                    "all_tensor_core_launches": {"tflops": fl / tm / 1e12 if tm > 0 else None,
                                                 "frac_of_bf16_sustained": fl / tm / 1e12 / peaks["bf16_sustained"] if tm > 0 else None,
                                                 "launches_per_step": sum(c["launches_per_step"] for c in tcs)}}
        return roofline, per_class

    def sub_record(self, name, steps, warmup):
        """Short measurement of another BASELINE.json configuration at the same N (no profiling, no CPU leg)."""
        from oracle import hifigan_oracle as O
        torch = self.torch
        step_device, step_e2e, info = self.setup(name)
        ms = self.timed(step_device, steps, warmup)
        ms_e2e = self.timed(step_e2e, max(2, steps // 2), 2)
        audio_s = self.world * O.audio_seconds(info["B"], info["T"])
        flops = O.forward_flops(info["cfg"], info["B"], info["T"]) * (3 if info["train"] else 1)
        rec = {"config": make_config(self.args, name, self.world), "value": audio_s / (ms * 1e-3), "unit": "audio-s/s",
               "ms_per_step": ms, "steps": steps, "warmup": warmup,
               "tflops_algorithmic_per_gpu": flops / (ms * 1e-3) / 1e12,
               "frac_of_bf16_peak_sustained": flops / (ms * 1e-3) / 1e12 / self.peaks["bf16_sustained"],
               "e2e": {"value": audio_s / (ms_e2e * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_e2e,
                       "h2d_bytes_per_step": info["h2d"], "d2h_bytes_per_step": info["d2h"]}}
        del step_device, step_e2e, info
        torch.cuda.empty_cache()
        return rec

    def mel_tail_record(self, B, T, steps=20, warmup=5):
        """the mel / STFT loss tail behind the decoder (vcd_mel_loss: loss + dy in one call)
        on the configs[1] segment shape, device-resident inputs; the CPU oracle (torch.stft autograd, fp32) beside it."""
        torch = self.torch
        from vcvits_b200 import mel as V
        kw = dict(n_fft=2048, num_mels=256, sampling_rate=48000, hop_size=512, win_size=2048, fmin=0.0, fmax=None)
        gen = torch.Generator(device="cpu").manual_seed(77)
        y_hat = (0.3 * torch.randn(B, 1, T, generator=gen)).clamp(-1, 1).to(self.dev)
        y_real = (0.3 * torch.randn(B, T, generator=gen)).clamp(-1, 1).to(self.dev)
        tgt = V.mel_spectrogram_torch(y_real, **kw)
        plan = V._plan(device=self.dev, **kw)

        def step():
            return plan.loss_and_grad(y_hat, tgt, 45.0)

        self.lib.vcd_launch_count(1)
        ms = self.timed(step, steps, warmup)
        launches = self.lib.vcd_launch_count(1) / (steps + warmup)
        plan.use_gemm_path(True)
        ms_gemm = self.timed(step, steps, warmup)
        plan.use_gemm_path(False)
        rows, nb = B * plan.frames(T), kw["n_fft"] // 2 + 1
        flops = 2.0 * 2.0 * rows * (kw["n_fft"] * 2 * nb + nb * kw["num_mels"])       # 4 GEMMs: STFT, mel and their transposes
        # algorithmic bytes of the fused per-frame kernel + overlap-add: y read once, target read, dframe written and read, dy written
        abytes = 4.0 * (B * T + rows * kw["num_mels"] + 2 * rows * kw["n_fft"] + B * T)
        from oracle import mel_oracle as M
        cores = use_all_host_threads()
        yc, tc = y_hat[:, 0].cpu(), tgt.cpu()

        def cpu_step():
            yy = yc.clone().requires_grad_(True)
            (torch.nn.functional.l1_loss(M.log_mel(yy, **kw), tc) * 45.0).backward()

        cpu_step()
        t0 = time.perf_counter()
        for _ in range(5):
            cpu_step()
        t_cpu = (time.perf_counter() - t0) / 5
        return {"what": "c_mel * l1(logmel(y_hat), y_mel) + d/dy_hat (vits/light/vcvits.py:96-115)", "B": B, "samples": T,
                "ms_per_call": ms, "launches_per_call": launches, "dtype": "f32",
                "path": "one CTA per frame: shared-memory radix-2 FFT + banded filterbank, forward and backward fused",
                "algorithmic_bytes": abytes, "gbs_algorithmic": abytes / (ms * 1e-3) / 1e9,
                "frac_of_hbm_copy_peak": abytes / (ms * 1e-3) / 1e9 / self.peaks["hbm"],
                "dense_gemm_path": {"ms_per_call": ms_gemm, "gflop": flops / 1e9, "tflops_ffma": flops / (ms_gemm * 1e-3) / 1e12},
                "cpu_oracle": {"ms_per_call": t_cpu * 1e3, "cores": cores, "kind": "port (torch.stft rfft + autograd, fp32)"}}

    def run(self):
        from oracle import hifigan_oracle as O
        args, torch, dist, lib = self.args, self.torch, self.dist, self.lib
        rank, world = self.rank, self.world
        step_device, step_e2e, info = self.setup(args.workload)
        cfg, B, T, train, model = info["cfg"], info["B"], info["T"], info["train"], info["model"]

        sampler = ClockSampler(self.local)
        if rank == 0:
            sampler.start()
        lib.vcd_launch_count(1)
        ms = self.timed(step_device, args.steps, args.warmup)
        host_ms = self.host_ms
        launches_total = lib.vcd_launch_count(1)
        clocks = sampler.stop() if rank == 0 else None
        launches_per_step = launches_total / (args.steps + args.warmup)

        e2e = None
        if not args.no_e2e:
            ms_e2e = self.timed(step_e2e, args.steps, max(3, args.warmup // 2))
            e2e = {"value": world * O.audio_seconds(B, T) / (ms_e2e * 1e-3), "unit": "audio-s/s",
                   "h2d_bytes_per_step": info["h2d"], "d2h_bytes_per_step": info["d2h"], "ms_per_step": ms_e2e}

        extra = {}
        if not args.no_extra and args.workload == "train_base_b16" and args.mode == "bf16":
            extra["train_48k_b32"] = self.sub_record("train_48k_b32", 10, 3)
            extra["infer_10s"] = self.sub_record("infer_10s", 5, 3)
            if rank == 0 and world == 1:
                extra["mel_loss_tail"] = self.mel_tail_record(16, 16384)

        # ---- rooflines, measured live with CUDA events around every launch (rank 0) ----
        roofline, classes, per_class = None, None, None
        if world > 1:
            dist.barrier()
        if rank == 0:
            classes = self.profile_classes(step_device, model)
            roofline, per_class = self.rooflines(classes, args.workload)
            if args.profile_classes:
                for c in classes:
                    print(json.dumps(c), file=sys.stderr)

        cpu_baseline = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cores = use_all_host_threads()
            cpu_mod, sB = cpu_bounded_sample(cfg, B, T, train, 3, 1, budget_s=20.0)
            t_cpu = cpu_time_steps(cpu_mod, cfg, sB, T, train, 3, 1)
            cpu_baseline = {"value": O.audio_seconds(sB, T) / t_cpu, "unit": "audio-s/s", "cores": cores, "kind": "port",
                            "sample": f"{'fwd+bwd' if train else 'fwd'} of {info['cfg_name']}, {sB} of the {B} utterances per step, "
                                      f"T={T}, fp32 CPU oracle, 3 timed steps after 1 warm-up ({t_cpu:.2f} s/step), {cores} threads"}

        if rank == 0:
            audio_s = world * O.audio_seconds(B, T)
            flops = O.forward_flops(cfg, B, T) * (3 if train else 1)
            value = audio_s / (ms * 1e-3)
            line = {"metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": args.mode, "data": "synthetic",
                    "config": make_config(args, args.workload, world),
                    "tflops_algorithmic": flops / (ms * 1e-3) / 1e12,
                    "frac_of_bf16_peak_sustained": flops / (ms * 1e-3) / 1e12 / self.peaks["bf16_sustained"],
                    "host_enqueue_ms_per_step": host_ms,
                    "gpu_launches": round(launches_per_step * args.steps), "gpu_launches_per_step": launches_per_step,
                    "clocks": clocks, "e2e": e2e, "roofline": roofline, "roofline_classes": per_class,
                    "cpu_baseline": cpu_baseline, "kernel_classes": classes, "configs": extra or None}
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
        if args.workload == "full_synth" and args.impl == "b200":
            from tools import full_synth
            full_synth.main(args)
        elif args.impl == "reference":
            run_reference(args)
        else:
            Bench(args).run()
    finally:
        builtins.print = _print
        real_stdout.flush()


if __name__ == "__main__":
    main()
