#!/usr/bin/env python
"""Headline benchmark: separated audio-seconds per second of TFLocoformerMSS on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16|fp32]

Workload (BASELINE.json configs[1], "Variant D" of SURVEY.md F4): n_fft 2048, hop 1024, 6 layers,
emb_dim 128, 4 heads, macaron ConvSwiGLU [384, 384], 4 sources; one step = one batch of 6-s 44.1 kHz
mono segments (stereo mixture averaged to mono, as every reference caller does) through
forward(mixture) -> dict of sources.  Random-init weights, synthetic mixtures.  N > 1: one process
per GPU (torchrun), each rank separates its own batch every step (segment sharding, no data-path
collective), time = max over ranks.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 44100
SEG = 264600  # 6 s
VARIANT_D = dict(n_fft=2048, hop_length=1024, n_sources=4, n_layers=6, emb_dim=128, norm_type="rmsgroupnorm",
                 num_groups=4, tf_order="ft", n_heads=4, flash_attention=True, attention_dim=128, pos_enc="rope",
                 ffn_type=["swiglu_conv1d", "swiglu_conv1d"], ffn_hidden_dim=[384, 384], conv1d_kernel=4,
                 conv1d_shift=1, dropout=0.0, eps=1e-5)
METRIC = "separated audio-sec/sec"
WORKLOAD = "musdb18 Variant D (n_fft 2048, hop 1024, 6 layers, emb 128, macaron 384) 6-s segments"


def algorithmic_flops(cfg, batch, n_samples):
    """SURVEY.md section 8d."""
    L, C, A, K = cfg["n_layers"], cfg["emb_dim"], cfg["attention_dim"], cfg["conv1d_kernel"]
    hid = cfg["ffn_hidden_dim"]
    hids = hid if isinstance(hid, list) else [hid]
    Tf, F = 1 + n_samples // cfg["hop_length"], cfg["n_fft"] // 2 + 1
    N = batch * Tf * F
    rows = batch * Tf * (F + K - 1) + batch * F * (Tf + K - 1)
    ffn = sum(L * 6 * K * C * h * rows for h in hids)
    sdpa = L * 4 * A * N * (F + Tf)
    proj = L * 2 * 8 * C * A * N
    return dict(ffn=ffn, sdpa=sdpa, proj=proj, total=ffn + sdpa + proj + 36 * C * N * (1 + cfg["n_sources"]))


def ffn_call_flops(cfg, batch, n_samples, axis, hidden):
    C, K = cfg["emb_dim"], cfg["conv1d_kernel"]
    Tf, F = 1 + n_samples // cfg["hop_length"], cfg["n_fft"] // 2 + 1
    rows = batch * Tf * (F + K - 1) if axis == 0 else batch * F * (Tf + K - 1)
    return 6 * K * C * hidden * rows


def make_mixture(batch, n_samples, seed=1234):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples) / SR
    x = 0.1 * torch.randn(batch, 2, n_samples, generator=g)
    for f0 in (55.0, 220.0, 880.0, 3520.0, 7040.0):
        x = x + 0.05 * torch.sin(2 * math.pi * f0 * t)
    return x.clamp(-1, 1).mean(1)  # stereo -> mono, inference/separate.py:135-139


def make_state_dict(cfg, seed=0):
    """Random-init weights of the named architecture (same init order as the reference constructor)."""
    import mss_tf_locoformer_b200 as pkg
    torch.manual_seed(seed)
    model = pkg.TFLocoformerMSS(**cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.ndim == 1 and not n.endswith("rope.freqs"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return model


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_forward(cfg, sd, mix, threads):
    import oracle
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    out = oracle.mss_forward(sd, cfg, mix)
    return out, time.perf_counter() - t0


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm (oracle port; the reference is PyTorch-on-CPU here and its RoPE
    dependency is not installable) on the host cores, bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cfg = dict(VARIANT_D)
    model = make_state_dict(cfg)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    budget = 150.0 / max(1, args.steps + args.warmup)          # seconds of CPU per step
    est_rate = 0.2 * min(1.0, cores / 8.0)                      # x real time measured on 8 EPYC cores
    audio_s = min(6.0, max(0.25, budget * est_rate))
    n = int(audio_s * SR)
    mix = make_mixture(1, n)
    for _ in range(args.warmup):
        cpu_oracle_forward(cfg, sd, mix, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_forward(cfg, sd, mix, cores)
    dt = time.perf_counter() - t0
    value = args.steps * n / SR / dt
    sample = f"{n / SR:.2f}-s mono segment, batch 1, fp32, {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=8, help="6-s segments per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--track", type=float, default=0.0,
                    help="BASELINE config 3: separate ONE synthetic track of this many seconds, 6-s segments at 50 %% overlap "
                         "sharded over the ranks (strong scaling); prints its own JSON line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    import mss_tf_locoformer_b200 as pkg
    from mss_tf_locoformer_b200 import _lib

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    cfg = dict(VARIANT_D)
    model = make_state_dict(cfg).to(dev)
    model.precision = args.precision
    B = args.batch
    mix_host = make_mixture(B, SEG, seed=1234 + rank).pin_memory()
    mix_dev = mix_host.to(dev)
    names = ["vocals", "drums", "bass", "other"]
    out_host = torch.empty((4, B, SEG), dtype=torch.float32).pin_memory()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        with torch.no_grad():
            return model(mix_dev)

    def step_e2e():
        with torch.no_grad():
            x = mix_host.to(dev, non_blocking=True)
            out = model(x)
            out_host.copy_(torch.stack([out[k] for k in names]), non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller needs the result on the host

    if args.track > 0:
        from mss_tf_locoformer_b200.segments import separate_track, segment_starts
        n_track = int(args.track * SR)
        track = make_mixture(1, n_track, seed=99)[0].to(dev)

        def step_track():
            with torch.no_grad():
                return separate_track(model, track, seg_len=SEG, batch=B)

        for _ in range(max(1, args.warmup // 3)):
            step_track()
        ms = timed(step_track, args.steps)
        if rank == 0:
            print(json.dumps({
                "metric": METRIC, "value": args.track * args.steps / (ms / 1e3), "unit": "audio-s/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
                "data": "synthetic",
                "config": {"workload": f"full-track {args.track:.0f}-s mono, {len(segment_starts(n_track, SEG))} segments of 6 s at "
                                       f"50 % overlap, sharded over {world} rank(s), NCCL sum-reduce stitch",
                           "batch_per_gpu": B, "precision": args.precision}}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    for _ in range(args.warmup):
        step_resident()
    n0 = lib.tfl_launch_count()
    with ClockSampler(local) as clocks:
        ms = timed(step_resident, args.steps)
    launches = lib.tfl_launch_count() - n0
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    audio_s = world * B * SEG / SR
    value = audio_s * args.steps / (ms / 1e3)
    e2e_value = audio_s * args.steps / (ms_e2e / 1e3)

    line = {
        "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "segment_samples": SEG, "precision": args.precision,
                   "l2": "activations per step (>= 1 GB) exceed the 126 MB L2; no explicit flush",
                   "x_realtime_per_gpu": value / world},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": world * mix_host.numel() * 4,
                "d2h_bytes_per_step": world * out_host.numel() * 4},
        "gpu_launches": int(launches) * world,
    }
    if rank == 0:
        line["clocks"] = clocks.summary()
        fl = algorithmic_flops(cfg, B, SEG)
        line["config"]["tflops_total_algorithmic"] = fl["total"] * world * args.steps / (ms / 1e3) / 1e12
        # ---- roofline of the dominant kernel: the ConvSwiGLU FFN (83 % of FLOPs), frequency axis ----
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        # the kernel below is timed alone (10 back-to-back launches): the burst cuBLAS figure is its denominator; the
        # whole-step fraction uses the sustained one (B200_PROFILING.md)
        peak = peaks.get("bf16_tflops", peaks.get("bf16_tflops_sustained", 1400.0))
        peak_sustained = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peaks else "fallback 1.4 PF"
        line["config"]["step_frac_of_sustained_bf16_peak"] = line["config"]["tflops_total_algorithmic"] / peak_sustained
        eng = model._ready()
        prec = 1 if args.precision == "bf16" else 0
        Tf, F = 1 + SEG // cfg["hop_length"], cfg["n_fft"] // 2 + 1
        x = torch.randn(B, Tf, F, cfg["emb_dim"], device=dev)
        for _ in range(3):
            eng.ffn_(0, 0, 0, x, prec)
        reps = 10
        l0 = lib.tfl_launch_count()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.ffn_(0, 0, 0, x, prec)
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / reps
        k_launches = (lib.tfl_launch_count() - l0) // reps
        k_flops = ffn_call_flops(cfg, B, SEG, 0, 384)
        achieved = k_flops / (k_ms / 1e3) / 1e12
        traffic = None   # DRAM bytes per launch from the committed ncu --set full capture of this very launch shape
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_ffn2_ncu_b8.json")))
            if B == 8 and args.precision == "bf16":
                traffic = cap["traffic_bytes_per_launch"]
        except (OSError, KeyError, ValueError):
            pass
        line["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak, "frac_of_sustained": achieved / peak_sustained, "traffic": traffic,
                            "kernel": "conv_swiglu_ffn (freq axis)",
                            "launches_per_call": int(k_launches), "ms_per_call": k_ms, "peak_source": peak_src,
                            "flops_per_call": k_flops}
        del x
        if not args.no_cpu_baseline:
            # ---- CPU baseline: the oracle port on the host cores, one 1.5-s segment (~10-20 s of CPU) ----
            cores = os.cpu_count() or 1
            n = SEG // 4
            sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
            mix = make_mixture(1, n)
            want, dt = cpu_oracle_forward(cfg, sd, mix, cores)
            with torch.no_grad():
                got = model(mix.to(dev))
            import oracle
            worst = min(oracle.si_sdr_db(got[k].cpu(), want[k]) for k in want)
            err = max(float((got[k].cpu() - want[k]).abs().max()) for k in want)
            line["cpu_baseline"] = {"value": n / SR / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                    "sample": f"one {n / SR:.2f}-s mono segment, batch 1, fp32 oracle, {dt:.1f} s"}
            line["parity"] = {"vs": "cpu oracle, same weights and input", "worst_si_sdr_db": worst, "max_abs": err}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
