#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 FAD hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: the CPU port of the reference

Metric (BASELINE.json): clips/sec for VGGish FAD embed + stats (+ Frechet).
Workload: BASELINE configs[3], VGGish part — 2 x 50 000 synthetic 10 s 16 kHz mono clips sharded over
8 GPUs = 2 x 6250 clips per GPU; with N GPUs the job is 2 x 6250 N clips (weak scaling).  One step =
one pass of the whole path over that batch: PCM -> log-mel patches -> VGGishCore -> fp64 {n, sum x,
sum x x^T} -> (one NCCL all-reduce) -> mean/cov -> Frechet distance.

  value : whole-job clips/s with the PCM already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e   : same job through the public API `FrechetAudioDistance.score_clips` from PINNED HOST buffers —
          chunked H2D copies and the D2H read of the FAD scalar are inside the timed region
  roofline     : the tcgen05 implicit-GEMM kernel (all 8 tensor-core layers): algorithmic FLOPs / summed
                 per-launch CUDA-event durations, against the measured sustained bf16 peak
  frontend     : the fused front-end (+ conv1) kernel: SURVEY 8d bytes per clip / its launch time, against the HBM peak
  e2e_pcm16    : like e2e, from pinned raw int16 PCM (the reference's dtype="int16" path; half the host -> device bytes)
  cpu_baseline : the oracle port of the reference's per-clip CPU path on a bounded sample (rank 0, N = 1)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CLIP_SAMPLES = 160000            # 10 s @ 16 kHz
CLIPS_PER_SET_PER_GPU = 6250     # BASELINE configs[3]: 2 x 50k clips over 8 GPUs
VGGISH_GFLOP_PER_CLIP = 17.278
METRIC = "clips/sec for VGGish FAD embed+stats (PCM -> log-mel -> VGGish -> mean/cov -> Frechet)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return {"bf16_tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                "bf16_tflops_burst": float(d.get("bf16_tflops", 0.0)) or None,
                "hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"bf16_tflops": 1400.0, "bf16_tflops_burst": None, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ synthetic data
def gen_clips_gpu(set_id: int, first: int, count: int, device, sr: int = 16000):
    """Same distributions as oracle/synth.py (background: white noise x U(0.02,0.3); eval: 1/f^alpha noise
    + 0.1 sine), generated with torch's Philox on the GPU, keyed by (set, clip index) so shards are
    invariant to the number of ranks."""
    import torch
    out = torch.empty((count, CLIP_SAMPLES), dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    blk = 250
    for b0 in range(0, count, blk):
        nb = min(blk, count - b0)
        g.manual_seed(0xFAD0 * 7919 + set_id * 1000003 + (first + b0))
        if set_id == 0:
            amp = torch.rand((nb, 1), generator=g, device=device) * 0.28 + 0.02
            x = torch.randn((nb, CLIP_SAMPLES), generator=g, device=device) * amp
        else:
            alpha = torch.rand((nb, 1), generator=g, device=device) * 0.7 + 0.3
            amp = torch.rand((nb, 1), generator=g, device=device) * 0.45 + 0.05
            f0 = torch.rand((nb, 1), generator=g, device=device) * 3900.0 + 100.0
            w = torch.randn((nb, CLIP_SAMPLES), generator=g, device=device)
            spec = torch.fft.rfft(w)
            f = torch.arange(spec.shape[1], device=device, dtype=torch.float32).clamp_(min=1.0)
            spec = spec * f[None, :] ** (-alpha / 2.0)
            x = torch.fft.irfft(spec, n=CLIP_SAMPLES)
            x = x / x.abs().amax(dim=1, keepdim=True) * amp
            t = torch.arange(CLIP_SAMPLES, device=device, dtype=torch.float32) / sr
            x = x + 0.1 * torch.sin(2 * torch.pi * f0 * t[None, :])
        out[b0:b0 + nb] = x.clamp_(-1.0, 1.0)
    return out


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_oracle_rate(n_bg: int, n_ev: int, warm: bool = True):
    """Times the oracle port of the reference's CPU path (per-clip loop, fad.py:302-408 + stats + Frechet)
    on a bounded sample with all host threads.  Returns (clips/s, seconds, fad, threads)."""
    import torch
    from oracle import networks, pipeline, synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = networks.vggish_random_state_dict(seed=0)
    ora = pipeline.OracleFAD("vggish", sd)
    bg = [synth.background_clip(i, CLIP_SAMPLES) for i in range(n_bg)]
    ev = [synth.eval_clip(i, CLIP_SAMPLES, 16000) for i in range(n_ev)]
    if warm:
        ora.get_embeddings(bg[:1])                    # one warm-up clip excluded
    t0 = time.perf_counter()
    fad, _, _ = ora.fad_from_clips(bg, ev)
    dt = time.perf_counter() - t0
    return (n_bg + n_ev) / dt, dt, float(fad), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_half = 8                                        # bounded sample per step: 8 + 8 ten-second clips
    for _ in range(max(args.warmup, 1)):
        cpu_oracle_rate(1, 1, warm=False)
    total = 0.0
    for _ in range(args.steps):
        r, dt, fad, threads = cpu_oracle_rate(n_half, n_half, warm=False)
        total += dt                                   # the path itself; synthesising the clips is not timed
    value = args.steps * 2 * n_half / total
    sample = f"{n_half}+{n_half} synthetic 10 s 16 kHz clips per step (per-clip loop like fad.py:317), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (front end f64)",
        "data": "synthetic", "config": workload_config(args.gpus, reference=True),
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus: int, reference: bool = False, clips_per_set: int = CLIPS_PER_SET_PER_GPU):
    return {
        "workload": "BASELINE configs[3] VGGish part: 2x50k synthetic 10 s 16 kHz mono clips over 8 GPUs = "
                    f"2x{clips_per_set} clips per GPU; this run: {n_gpus} GPU(s), 2x{clips_per_set * n_gpus} clips"
                    + (" (reference arm: bounded sample of the same clips per step)" if reference else ""),
        "clips_per_set_per_gpu": clips_per_set, "clip_seconds": 10, "sample_rate": 16000,
        "weights": "random-init VGGishCore (seed 0, He-normal), identical in both arms",
        "parallelism": f"clip-sharded dp{n_gpus}, one fp64 all-reduce of (n, sum x, sum x x^T)",
        "l2": "inputs (8 GB PCM per GPU per step) are far larger than the 126 MB L2; no explicit flush",
    }


# ------------------------------------------------------------------------------------------------ our arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from frechet_audio_distance_exported_b200 import FrechetAudioDistance
    from frechet_audio_distance_exported_b200.dist import shard_bounds
    from oracle import networks                        # weights only: the same seeded state_dict as the CPU arm

    sd = networks.vggish_random_state_dict(seed=0)
    fad = FrechetAudioDistance(model_name="vggish", state_dict=sd, precision=args.precision)
    eng = fad.engine
    n_set = args.clips_per_set * world                 # whole job, per set
    lo, hi = shard_bounds(n_set, rank, world)
    bg = gen_clips_gpu(0, lo, hi - lo, dev)
    ev = gen_clips_gpu(1, lo, hi - lo, dev)
    torch.cuda.synchronize()
    d = eng.dim
    acc = torch.zeros(2 * (1 + d + d * d), dtype=torch.float64, device=dev)
    half = acc.numel() // 2

    def step_device():
        acc.zero_()
        eng.stats_accumulate(eng.embed_pcm(bg), acc[:half])
        eng.stats_accumulate(eng.embed_pcm(ev), acc[half:])
        eng.allreduce_acc(acc)
        mu1, s1 = eng.stats_finalize(acc[:half], d)
        mu2, s2 = eng.stats_finalize(acc[half:], d)
        return eng.frechet(mu1, s1, mu2, s2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    ms_total, out = timed(step_device, args.steps)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    fad_value = float(out[0].item())
    ms_step = ms_total / args.steps
    value = 2 * n_set / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel: per-launch CUDA-event durations of the tcgen05 layers, one extra step
    eng.profile_enable(True)
    step_device()
    torch.cuda.synchronize()
    gemm_ms, gemm_flops, gemm_launches = eng.profile_read()
    front_ms = eng.front_ms
    eng.profile_enable(False)
    peaks = _peaks()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        # measured DRAM bytes per patch over the 8 layer launches (ncu --set full, 4000-patch chunk), scaled to the
        # patches the launches of this run process: average DRAM bytes per launch, like `achieved`
        if tj.get("gemm_dram_bytes_per_patch") and gemm_launches:
            traffic = tj["gemm_dram_bytes_per_patch"] * (2 * (hi - lo) * 10) / gemm_launches
        else:
            traffic = tj.get("gemm_dram_bytes_per_launch")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": traffic,
                "kernel": "fadb_gemm_tc_kernel (tcgen05 implicit GEMM: 5 conv3x3 + 3 FC layers)",
                "how": f"algorithmic 2*M*N*K FLOPs of {gemm_launches} launches / sum of their CUDA-event durations "
                       f"({gemm_ms:.2f} ms of a {ms_step:.2f} ms step); peak = {peaks['source']}",
                "step_share": gemm_ms / ms_step if ms_step > 0 else None,
                # the denominator is cuBLAS (torch.matmul 8192^3) run back to back under the same power cap; a
                # fraction above 1 means these launches ran faster than that loop, not faster than the silicon
                "peak_burst": peaks["bf16_tflops_burst"],
                "frac_of_burst": achieved / peaks["bf16_tflops_burst"] if peaks["bf16_tflops_burst"] else None}

    # ---- the front-end stage next to it (north_star: "reported as achieved HBM GB/s"): SURVEY 8d bytes per clip
    # (640 000 B of PCM read + 245 760 B of fp32 patches); the fused kernel also runs conv1, so it writes the
    # pooled conv1 activations (1 966 080 B per clip, bf16) instead of the patches
    clips_step = 2 * (hi - lo)
    frontend = {"bound": "hbm", "achieved": clips_step * 885760 / (front_ms * 1e-3) / 1e9 if front_ms > 0 else None,
                "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": clips_step * 885760 / (front_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if front_ms > 0 else None,
                "kernel": "fadb_vggish_front_conv1_tc_kernel (fp64 FFT log-mel front end fused with tcgen05 conv1)",
                "ms_per_step": front_ms, "step_share": front_ms / ms_step if ms_step > 0 else None,
                "actual_bytes_per_clip": 640000 + 10 * 48 * 32 * 64 * 2,
                "note": "not HBM bound: a latency / shared-memory limited fp64 FFT phase plus the drain of the conv1 accumulators"}

    # ---- e2e through the public API from pinned host memory
    bg_h = torch.empty(bg.shape, dtype=torch.float32).pin_memory()
    ev_h = torch.empty(ev.shape, dtype=torch.float32).pin_memory()
    bg_h.copy_(bg); ev_h.copy_(ev)
    torch.cuda.synchronize()
    fad.process_group = None
    # what the host -> device link gives on this box for the same pinned buffer (explains how close e2e is to it)
    probe = torch.empty_like(bg)
    probe.copy_(bg_h, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    probe.copy_(bg_h, non_blocking=True); torch.cuda.synchronize()
    h2d_gbs = bg_h.numel() * 4 / (time.perf_counter() - t0) / 1e9
    del probe
    e2e_steps = max(2, min(args.steps, 3))
    fad.score_clips(bg_h, ev_h)                        # warm-up (allocates the double buffers)
    barrier()
    t0 = time.perf_counter()
    e2e_fad = None
    for _ in range(e2e_steps):
        e2e_fad = fad.score_clips(bg_h, ev_h)          # includes H2D of every clip and D2H of the scalar
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_s = torch.tensor([(t1 - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = 2 * n_set / float(e2e_s.item())
    h2d = int(bg_h.numel() + ev_h.numel()) * 4
    e2e = {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
           "api": "FrechetAudioDistance.score_clips(pinned_host_bg, pinned_host_ev)", "steps": e2e_steps,
           "fad": e2e_fad, "h2d_gbs_needed": h2d / (float(e2e_s.item())) / 1e9, "h2d_gbs_link_alone": h2d_gbs}

    # ---- same job from raw 16-bit PCM (the WAV sample format; reference dtype="int16", fad.py:145-149): half the bytes
    q = lambda t: (t * 32767.0).round_().to(torch.int16)
    bg16 = torch.empty(bg.shape, dtype=torch.int16).pin_memory()
    ev16 = torch.empty(ev.shape, dtype=torch.int16).pin_memory()
    for src, dst in ((bg, bg16), (ev, ev16)):
        for c0 in range(0, src.shape[0], 512):
            dst[c0:c0 + 512].copy_(q(src[c0:c0 + 512].clone()))
    torch.cuda.synchronize()
    fad.score_clips(bg16, ev16)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e16_fad = fad.score_clips(bg16, ev16)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e16_s = torch.tensor([(t1 - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e16_s, op=dist.ReduceOp.MAX)
    e2e_pcm16 = {"value": 2 * n_set / float(e16_s.item()), "unit": "clips/s",
                 "h2d_bytes_per_step": int(bg16.numel() + ev16.numel()) * 2, "d2h_bytes_per_step": 8,
                 "api": "FrechetAudioDistance.score_clips(pinned int16 PCM)", "steps": e2e_steps, "fad": e2e16_fad}
    del bg16, ev16

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r, dt, fad_cpu, threads = cpu_oracle_rate(args.cpu_clips, args.cpu_clips)
        # parity spot check on the very same clips
        from oracle import synth
        import numpy as np
        cb = torch.from_numpy(np.stack([synth.background_clip(i, CLIP_SAMPLES) for i in range(args.cpu_clips)]))
        ce = torch.from_numpy(np.stack([synth.eval_clip(i, CLIP_SAMPLES, 16000) for i in range(args.cpu_clips)]))
        fad_gpu_same = fad.score_clips(cb, ce)
        eng.set_precision("bf16x3")                    # the parity mode (split-bf16 + exact accumulation) on the same clips
        fad_gpu_x3 = fad.score_clips(cb, ce)
        eng.set_precision(args.precision)
        cpu = {"value": r, "unit": "clips/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_clips}+{args.cpu_clips} synthetic 10 s clips through oracle/pipeline.py "
                         f"(per-clip loop like fad.py:317, torch CPU fp32 + NumPy f64 front end), {dt:.1f} s",
               "fad_cpu": fad_cpu, "fad_gpu_same_clips": fad_gpu_same,
               "fad_rel_diff": abs(fad_gpu_same - fad_cpu) / abs(fad_cpu),
               "fad_gpu_same_clips_bf16x3": fad_gpu_x3, "fad_rel_diff_bf16x3": abs(fad_gpu_x3 - fad_cpu) / abs(fad_cpu)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (split-bf16, 3 MMAs/product)",
            "data": "synthetic (torch Philox on GPU, distributions of oracle/synth.py)",
            "config": workload_config(world, clips_per_set=args.clips_per_set),
            "fad": fad_value, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "frontend": frontend, "e2e_pcm16": e2e_pcm16, "cpu_baseline": cpu,
            "frac_of_tensor_roofline_whole_step": (value / world) * VGGISH_GFLOP_PER_CLIP / 1e3 / peaks["bf16_tflops"],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--clips-per-set", type=int, default=CLIPS_PER_SET_PER_GPU)
    ap.add_argument("--cpu-clips", type=int, default=64, help="clips per set of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
