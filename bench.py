#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 FAD hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: the CPU port of the reference

Metric (BASELINE.json): clips/sec for VGGish FAD embed + stats (+ Frechet); Frechet ms at d = 2048.
Workload: BASELINE configs[3], VGGish part — 2 x 50 000 synthetic 10 s 16 kHz mono clips sharded over
8 GPUs = 2 x 6250 clips per GPU; with N GPUs the job is 2 x 6250 N clips (weak scaling).  One step =
one pass of the whole path over that batch: PCM -> log-mel patches -> VGGishCore -> fp64 {n, sum x,
sum x x^T} -> (one NCCL all-reduce) -> mean/cov -> Frechet distance.

The timed precision is the library default "fp16x2" (fp16 activations x split-fp16 weights, 2 MMAs per product):
the one whose FAD is within the north-star 1e-4 of the reference (cpu_baseline.fad_rel_diff).  The single-pass
modes (bf16, fp16: half the tensor work, FAD within 1e-4 .. 1e-3) are timed beside it under `modes`.

  value : whole-job clips/s with the PCM already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e   : same job through the public API `FrechetAudioDistance.score_clips` from PINNED HOST buffers —
          chunked H2D copies and the D2H read of the FAD scalar are inside the timed region
  roofline     : the tcgen05 implicit-GEMM kernel (all 8 tensor-core layers): algorithmic FLOPs / summed
                 per-launch CUDA-event durations, against the measured sustained bf16 peak
  frontend     : the fused front-end (+ conv1) kernel: SURVEY 8d bytes per clip / its launch time, against the HBM peak
  e2e_pcm16    : like e2e, from pinned raw int16 PCM (the reference's dtype="int16" path; half the host -> device bytes)
  modes        : device-resident clips/s and FAD deviation of the other precision modes
  frechet_ms / stats_ms : BASELINE configs[4] — Frechet at d = 128 / 512 / 2048 (+ the N = 1000 < d case of configs[1])
                 and mean/cov at N = 1e5, beside scipy / numpy on the host
  models       : PANN-16k, PANN-32k, CLAP (CNN14) clips/s on 512 ten-second clips, their tensor roofline and the
                 embedding deviation against the CPU oracle on 4 clips
  shard_invariance_rel : |FAD(all ranks, sharded) - FAD(rank 0 alone)| / FAD on the same fixed 2 x 256 clips
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
GFLOP_PER_CLIP = {"vggish": 17.278, "pann-8k": 41.3627, "pann-16k": 41.3627, "pann-32k": 41.3627, "clap": 40.0843}
SAMPLE_RATE = {"vggish": 16000, "pann-8k": 8000, "pann-16k": 16000, "pann-32k": 32000, "clap": 48000}
FRONT_BYTES_PER_CLIP = {"vggish": 885760, "pann-8k": 584192, "pann-16k": 904192, "pann-32k": 1544192, "clap": 2176256}
METRIC = "clips/sec for VGGish FAD embed+stats (PCM -> log-mel -> VGGish -> mean/cov -> Frechet)"
DTYPE_NAMES = {
    "fp16x2": "fp16 activations x split-fp16 (hi+lo) weights, 2 MMAs per product, fp32 accumulate",
    "fp16": "fp16, fp32 accumulate", "bf16": "bf16, fp32 accumulate",
    "bf16x3": "bf16x3 (split-bf16, 3 MMAs per product, exact segment sums)",
}


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
def gen_clips_gpu(set_id: int, first: int, count: int, device, sr: int = 16000, n_samples: int = CLIP_SAMPLES):
    """Same distributions as oracle/synth.py (background: white noise x U(0.02,0.3); eval: 1/f^alpha noise
    + 0.1 sine), generated with torch's Philox on the GPU in blocks of 250 clips seeded by the block's first clip."""
    import torch
    out = torch.empty((count, n_samples), dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    blk = 250 if n_samples <= 160000 else 64
    for b0 in range(0, count, blk):
        nb = min(blk, count - b0)
        g.manual_seed(0xFAD0 * 7919 + set_id * 1000003 + (first + b0))
        if set_id == 0:
            amp = torch.rand((nb, 1), generator=g, device=device) * 0.28 + 0.02
            x = torch.randn((nb, n_samples), generator=g, device=device) * amp
        else:
            alpha = torch.rand((nb, 1), generator=g, device=device) * 0.7 + 0.3
            amp = torch.rand((nb, 1), generator=g, device=device) * 0.45 + 0.05
            f0 = torch.rand((nb, 1), generator=g, device=device) * 3900.0 + 100.0
            w = torch.randn((nb, n_samples), generator=g, device=device)
            spec = torch.fft.rfft(w)
            f = torch.arange(spec.shape[1], device=device, dtype=torch.float32).clamp_(min=1.0)
            spec = spec * f[None, :] ** (-alpha / 2.0)
            x = torch.fft.irfft(spec, n=n_samples)
            x = x / x.abs().amax(dim=1, keepdim=True) * amp
            t = torch.arange(n_samples, device=device, dtype=torch.float32) / sr
            x = x + 0.1 * torch.sin(2 * torch.pi * f0 * t[None, :])
        out[b0:b0 + nb] = x.clamp_(-1.0, 1.0)
    return out


# ------------------------------------------------------------------------------------------------ CPU legs
# The oracle port of the reference's CPU path (per-clip loop, fad.py:302-408 + stats + Frechet) is timed in a
# SUBPROCESS so that the thread configuration is clean.  Two configurations are calibrated on a small sample and
# the faster one is used and named in the result: torch's intra-op pool on every core with NumPy's BLAS pool
# (a) single-threaded (OMP_NUM_THREADS=1, what torchrun exports) or (b) left at its default.  Round 1 used (b) blindly
# and the two pools oversubscribed the cores (5-6x slower than (a) on the 16-32-core GPU boxes).
CPU_CONFIGS = {
    "omp1_torchN": {"OMP_NUM_THREADS": "1", "MKL_NUM_THREADS": "1", "OPENBLAS_NUM_THREADS": "1"},
    "default_torchN": {},
}


def _cpu_worker(spec: dict) -> None:
    """child process: time `steps` passes of the oracle port over n_bg + n_ev ten-second clips"""
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))          # undo any NUMA binding inherited from the parent
    except Exception:
        pass
    import torch
    from oracle import networks, pipeline, synth
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    sd = networks.vggish_random_state_dict(seed=0)
    ora = pipeline.OracleFAD("vggish", sd)
    bg = [synth.background_clip(i, CLIP_SAMPLES) for i in range(spec["n_bg"])]
    ev = [synth.eval_clip(i, CLIP_SAMPLES, 16000) for i in range(spec["n_ev"])]
    for _ in range(spec.get("warmup", 1)):
        ora.get_embeddings(bg[:1])                                    # warm-up clips are not timed
    times, fad = [], None
    for _ in range(spec["steps"]):
        t0 = time.perf_counter()
        fad, _, _ = ora.fad_from_clips(bg, ev)
        times.append(time.perf_counter() - t0)
    print("CPUWORKER " + json.dumps({"step_s": times, "fad": float(fad), "threads": threads,
                                     "omp": os.environ.get("OMP_NUM_THREADS", "unset")}), flush=True)


def _run_cpu_worker(cfg: str, n_bg: int, n_ev: int, steps: int, warmup: int = 1) -> dict:
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS")}
    env.update(CPU_CONFIGS[cfg])
    env["CUDA_VISIBLE_DEVICES"] = ""
    spec = json.dumps({"n_bg": n_bg, "n_ev": n_ev, "steps": steps, "warmup": warmup})
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-worker", spec], env=env, capture_output=True,
                       text=True, cwd=ROOT)
    for line in r.stdout.splitlines():
        if line.startswith("CPUWORKER "):
            return json.loads(line[len("CPUWORKER "):])
    raise RuntimeError("CPU worker failed: " + r.stderr[-2000:])


def cpu_oracle_rate(n_bg: int, n_ev: int, steps: int = 1):
    """-> dict(value clips/s, seconds per step, fad, threads, config, calibration)"""
    calib = {}
    for cfg in CPU_CONFIGS:
        r = _run_cpu_worker(cfg, 3, 3, 1)
        calib[cfg] = 6.0 / r["step_s"][0]
    best = max(calib, key=calib.get)
    r = _run_cpu_worker(best, n_bg, n_ev, steps)
    total = sum(r["step_s"])
    return {"value": steps * (n_bg + n_ev) / total, "s_per_step": total / steps, "fad": r["fad"], "threads": r["threads"],
            "config": best, "calibration_clips_per_s": calib}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_half = args.cpu_clips_ref                       # bounded sample per step
    r = cpu_oracle_rate(n_half, n_half, args.steps)
    value = r["value"]
    sample = (f"{n_half}+{n_half} synthetic 10 s 16 kHz clips per step (per-clip loop like fad.py:317), {args.steps} steps, "
              f"thread config {r['config']} (calibrated: {r['calibration_clips_per_s']})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["s_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (front end f64)",
        "data": "synthetic", "config": workload_config(args.gpus, reference=True),
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": r["threads"], "kind": "port", "sample": sample,
                         "thread_config": r["config"], "fad": r["fad"]},
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


# ------------------------------------------------------------------------------------------------ extras (N = 1)
def _gpu_ms(fn, reps=3):
    import torch
    fn(); torch.cuda.synchronize()
    best, out = None, None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        best = t if best is None or t < best else best
    return best, out


def stats_frechet_microbench(eng, cpu: bool):
    """BASELINE configs[4]: mean/cov at N = 1e5 and the Frechet kernel chain at d = 128 / 512 / 2048 (+ the
    rank-deficient N = 1000 < d = 2048 case of configs[1]); x = z A + b as in SURVEY 8d.  CPU: np.cov + scipy sqrtm
    (the reference's fad.py:494-495,535-555) once per d."""
    import numpy as np
    import torch
    from oracle import stats as ostats
    dev = eng.device

    def make(n, d, seed, scale, shift):
        g = torch.Generator(device=dev).manual_seed(seed)
        a = torch.randn(d, d, generator=g, device=dev) / d ** 0.5
        out = torch.empty(n, d, device=dev)
        for i in range(0, n, 65536):
            m = min(65536, n - i)
            out[i:i + m] = (torch.randn(m, d, generator=g, device=dev) @ a) * scale + shift
        return out

    frechet, stats = {}, {}
    for d, n in ((128, 100_000), (512, 100_000), (2048, 100_000), (2048, 1000)):
        x1, x2 = make(n, d, 1, 1.0, 0.0), make(n, d, 2, 1.1, 0.05)

        def stats_both():
            out = []
            for x in (x1, x2):
                acc = eng.new_acc(d)
                eng.stats_accumulate(x, acc)
                out.append(eng.stats_finalize(acc, d))
            return out
        ms_stats, ((mu1, s1), (mu2, s2)) = _gpu_ms(stats_both)
        ms_fr, fr = _gpu_ms(lambda: eng.frechet(mu1, s1, mu2, s2))
        key = f"d{d}" if n != 1000 else f"d{d}_n1000_rank_deficient"
        frechet[key] = {"gpu_ms": ms_fr, "n_per_set": n, "fad": float(fr[0])}
        if n != 1000:
            stats[key] = {"gpu_ms_both_sets": ms_stats, "n_per_set": n, "read_gbs": 2 * n * d * 4 / ms_stats / 1e6,
                          "flops_per_s_T": 2 * n * d * (d + 1) / ms_stats / 1e9, "kernel": "fp64 DFMA syrk (default)"}
            if d >= 512:
                # opt-in tensor-core syrk (split-fp16 tcgen05 GEMM, fadb_set_tensor_syrk): speed and what it costs in accuracy
                eng.set_tensor_syrk(True)
                ms_tc, ((_, t1), (_, t2)) = _gpu_ms(stats_both)
                eng.set_tensor_syrk(False)
                frt = eng.frechet(mu1, t1, mu2, t2)
                stats[key]["tensor_core"] = {
                    "gpu_ms_both_sets": ms_tc, "flops_per_s_T": 2 * n * d * (d + 1) / ms_tc / 1e9,
                    "sigma_rel_err_vs_fp64_kernel": float((t1 - s1).abs().max() / s1.abs().max()),
                    "fad_rel_err_vs_fp64_kernel": abs(float(frt[0]) - float(fr[0])) / abs(float(fr[0]))}
        if cpu:
            h1, h2 = x1[:20000].cpu().numpy(), x2[:20000].cpu().numpy()       # bounded CPU sample
            t0 = time.perf_counter()
            m1, c1 = ostats.embd_statistics(h1)
            m2, c2 = ostats.embd_statistics(h2)
            t_stats = time.perf_counter() - t0
            if n != 1000:
                stats[key]["cpu_numpy_ms_both_sets_per_1e5_rows"] = t_stats * 1e3 * (n / h1.shape[0])
            # Frechet on the GPU's own statistics (identical input to both): scipy sqrtm path of the reference
            mu1h, s1h, mu2h, s2h = (t.cpu().numpy() for t in (mu1, s1, mu2, s2))
            t0 = time.perf_counter()
            ref = ostats.frechet_distance(mu1h, s1h, mu2h, s2h)
            frechet[key]["cpu_scipy_ms"] = (time.perf_counter() - t0) * 1e3
            frechet[key]["rel_diff_vs_scipy"] = abs(float(fr[0]) - float(ref)) / abs(float(ref))
        del x1, x2
        torch.cuda.empty_cache()
    return frechet, stats


def model_throughput(name: str, n_clips: int, peaks, cpu: bool):
    """PANN / CLAP (CNN14) on `n_clips` ten-second clips, device-resident: clips/s through PCM -> log-mel -> CNN14 ->
    statistics, the tensor roofline of its GEMM launches, and the embedding deviation against the CPU oracle."""
    import numpy as np
    import torch
    from frechet_audio_distance_exported_b200 import Engine
    from oracle import networks, pipeline, synth
    sr = SAMPLE_RATE[name]
    sd = networks.cnn14_random_state_dict(seed=1, clap_head=(name == "clap"))
    eng = Engine(name, sd)
    pcm = gen_clips_gpu(1, 0, n_clips, eng.device, sr=sr, n_samples=10 * sr)
    acc = eng.new_acc()

    def step():
        acc.zero_()
        eng.stats_accumulate(eng.embed_pcm(pcm), acc)
    res = {"clips": n_clips, "clip_seconds": 10, "sample_rate": sr, "dtype": eng.precision}
    for prec in ("fp16x2", "bf16"):
        eng.set_precision(prec)
        ms, _ = _gpu_ms(step, reps=2)
        eng.profile_enable(True)
        step()
        torch.cuda.synchronize()
        gemm_ms, gemm_flops, launches = eng.profile_read()
        front_ms = eng.front_ms
        eng.profile_enable(False)
        ach = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        passes = 2 if prec == "fp16x2" else 1
        r = {"clips_per_s": n_clips / ms * 1e3, "ms": ms,
             "roofline": {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                          "frac": ach / peaks["bf16_tflops"], "mma_passes": passes, "executed": ach * passes,
                          "frac_executed": ach * passes / peaks["bf16_tflops"], "launches": launches,
                          "step_share": gemm_ms / ms},
             "frontend_gbs": n_clips * FRONT_BYTES_PER_CLIP[name] / (front_ms * 1e-3) / 1e9 if front_ms > 0 else None,
             "whole_step_tflops": n_clips * GFLOP_PER_CLIP[name] / ms}
        if prec == "fp16x2":
            res.update(r)
        else:
            res["bf16"] = r
    if cpu:
        eng.set_precision("fp16x2")
        k = 4
        clips = [synth.eval_clip(300 + i, 10 * sr, sr) for i in range(k)]
        ref = pipeline.OracleFAD(name, sd).get_embeddings(clips)
        out = eng.embed_pcm(torch.from_numpy(np.stack(clips)).to(eng.device)).cpu().numpy()
        res["embedding_rel_max_err_vs_oracle_4_clips"] = float(np.abs(out - ref).max() / np.abs(ref).max())
        eng.set_precision("bf16")
        out = eng.embed_pcm(torch.from_numpy(np.stack(clips)).to(eng.device)).cpu().numpy()
        res["bf16"]["embedding_rel_max_err_vs_oracle_4_clips"] = float(np.abs(out - ref).max() / np.abs(ref).max())
    del eng, pcm
    torch.cuda.empty_cache()
    return res


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
    from frechet_audio_distance_exported_b200 import FrechetAudioDistance
    from frechet_audio_distance_exported_b200.dist import shard_bounds
    from frechet_audio_distance_exported_b200.stream import HostRing, bind_to_gpu_numa_node
    numa_cpus = None if args.no_numa else bind_to_gpu_numa_node(local)   # before any pinned allocation (first touch)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
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

    def wall_max(seconds: float) -> float:
        t = torch.tensor([seconds], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    passes = {"fp16x2": 2, "bf16x3": 3}.get(args.precision, 1)

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
        # measured DRAM bytes per patch over the 8 layer launches (ncu --set full), scaled to the patches the launches
        # of this run process: average DRAM bytes per launch, like `achieved`
        per_patch = tj.get(f"gemm_dram_bytes_per_patch_{args.precision}", tj.get("gemm_dram_bytes_per_patch"))
        if per_patch and gemm_launches:
            traffic = per_patch * (2 * (hi - lo) * 10) / gemm_launches
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": traffic,
                "kernel": "fadb_gemm_tc_kernel (tcgen05 implicit GEMM: 5 conv3x3 + 3 FC layers)",
                "how": f"ALGORITHMIC 2*M*N*K FLOPs of {gemm_launches} launches / sum of their CUDA-event durations "
                       f"({gemm_ms:.2f} ms of a {ms_step:.2f} ms step); peak = {peaks['source']}",
                "step_share": gemm_ms / ms_step if ms_step > 0 else None,
                # the timed precision issues `mma_passes` MMAs per algorithmic product (fp16x2: A*W_hi + A*W_lo), so the
                # tensor pipe executes `executed` TFLOP/s; `frac` stays algorithmic as the contract asks
                "mma_passes": passes, "executed": achieved * passes,
                "frac_executed": achieved * passes / peaks["bf16_tflops"],
                # the denominator is cuBLAS (torch.matmul 8192^3) run back to back under the same power cap; a
                # fraction above 1 means these launches ran faster than that loop, not faster than the silicon
                "peak_burst": peaks["bf16_tflops_burst"],
                "frac_executed_of_burst": achieved * passes / peaks["bf16_tflops_burst"] if peaks["bf16_tflops_burst"] else None}

    # ---- the front-end stage next to it (north_star: "reported as achieved HBM GB/s"): SURVEY 8d bytes per clip
    # (640 000 B of PCM read + 245 760 B of fp32 patches); the fused kernel also runs conv1, so it writes the
    # pooled conv1 activations (1 966 080 B per clip, 16-bit) instead of the patches
    clips_step = 2 * (hi - lo)
    # bytes the fused kernel really moves per clip: PCM in, 16-bit conv1 activations out, and in fp16x2 their W-padded
    # e4m3 copy (48 x 34 x 64 per patch) for conv2's low-order pass
    actual_clip = 640000 + 10 * 48 * 32 * 64 * 2 + (10 * 48 * 34 * 64 if args.precision == "fp16x2" and
                                                      os.environ.get("FADB_LO_FP8_C64", "1") != "0" and
                                                      os.environ.get("FADB_LO_FP8", "1") != "0" else 0)
    frontend = {"bound": "hbm", "achieved": clips_step * 885760 / (front_ms * 1e-3) / 1e9 if front_ms > 0 else None,
                "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": clips_step * 885760 / (front_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if front_ms > 0 else None,
                "kernel": "fadb_vggish_front_conv1_tc_kernel (fp64 FFT log-mel front end fused with tcgen05 conv1)",
                "ms_per_step": front_ms, "step_share": front_ms / ms_step if ms_step > 0 else None,
                "actual_bytes_per_clip": actual_clip,
                "actual_gbs": clips_step * actual_clip / (front_ms * 1e-3) / 1e9 if front_ms > 0 else None}

    # ---- the other precision modes, device-resident, 2 timed steps each (same data, same step function)
    modes = {args.precision: {"value": value, "ms_per_step": ms_step, "fad": fad_value, "mma_passes": passes}}
    for prec in ("bf16", "fp16", "bf16x3"):
        if prec == args.precision or args.no_modes:
            continue
        eng.set_precision(prec)
        step_device()
        ms_m, out_m = timed(step_device, 2)
        modes[prec] = {"value": 2 * n_set / (ms_m / 2 * 1e-3), "ms_per_step": ms_m / 2, "fad": float(out_m[0].item()),
                       "mma_passes": 3 if prec == "bf16x3" else 1,
                       "fad_rel_diff_vs_timed_mode": abs(float(out_m[0].item()) - fad_value) / abs(fad_value)}
    eng.set_precision(args.precision)

    # ---- e2e through the public API from pinned host memory
    bg_h = torch.empty(bg.shape, dtype=torch.float32).pin_memory()
    ev_h = torch.empty(ev.shape, dtype=torch.float32).pin_memory()
    bg_h.copy_(bg); ev_h.copy_(ev)
    torch.cuda.synchronize()
    fad.process_group = None
    # what the host -> device link gives on this box for the same pinned buffer, all ranks copying at the same time:
    # (1) one big copy, (2) the chunked ring the pipeline uses, with no kernels behind it
    probe = torch.empty_like(bg)
    probe.copy_(bg_h, non_blocking=True); barrier()
    t0 = time.perf_counter()
    probe.copy_(bg_h, non_blocking=True); torch.cuda.synchronize()
    h2d_gbs = bg_h.numel() * 4 / wall_max(time.perf_counter() - t0) / 1e9
    del probe
    ring = HostRing(dev, depth=4)
    ring.run(bg_h, 512, lambda dv, c0, nc: None); barrier()
    t0 = time.perf_counter()
    ring.run(bg_h, 512, lambda dv, c0, nc: None); torch.cuda.synchronize()
    h2d_ring_gbs = bg_h.numel() * 4 / wall_max(time.perf_counter() - t0) / 1e9
    del ring
    e2e_steps = max(2, min(args.steps, 3))
    fad.score_clips(bg_h, ev_h)                        # warm-up (allocates the staging ring)
    barrier()
    t0 = time.perf_counter()
    e2e_fad = None
    for _ in range(e2e_steps):
        e2e_fad = fad.score_clips(bg_h, ev_h)          # includes H2D of every clip and D2H of the scalar
    torch.cuda.synchronize()
    e2e_s = wall_max((time.perf_counter() - t0) / e2e_steps)
    e2e_value = 2 * n_set / e2e_s
    h2d = int(bg_h.numel() + ev_h.numel()) * 4
    e2e = {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
           "api": "FrechetAudioDistance.score_clips(pinned_host_bg, pinned_host_ev)", "steps": e2e_steps,
           "fad": e2e_fad, "h2d_gbs_needed": h2d / e2e_s / 1e9, "h2d_gbs_link_alone": h2d_gbs,
           "h2d_gbs_ring_no_kernels": h2d_ring_gbs, "numa_bound_cpus": numa_cpus,
           "device_resident_value": value, "frac_of_device_resident": e2e_value / value}
    if args.precision != "bf16" and not args.no_modes:          # the single-pass mode needs twice the PCM rate
        eng.set_precision("bf16")
        fad.score_clips(bg_h, ev_h); barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            fad.score_clips(bg_h, ev_h)
        torch.cuda.synchronize()
        s_b = wall_max((time.perf_counter() - t0) / 2)
        modes["bf16"]["e2e_value"] = 2 * n_set / s_b
        modes["bf16"]["e2e_h2d_gbs_needed"] = h2d / s_b / 1e9
        eng.set_precision(args.precision)

    # ---- same job from raw 16-bit PCM (the WAV sample format; reference dtype="int16", fad.py:145-149): half the bytes
    q = lambda t: (t * 32767.0).round_().to(torch.int16)
    bg16 = torch.empty(bg.shape, dtype=torch.int16).pin_memory()
    ev16 = torch.empty(ev.shape, dtype=torch.int16).pin_memory()
    for src, dst in ((bg, bg16), (ev, ev16)):
        for c0 in range(0, src.shape[0], 512):
            dst[c0:c0 + 512].copy_(q(src[c0:c0 + 512].clone()))
    torch.cuda.synchronize()
    del bg_h, ev_h
    fad.score_clips(bg16, ev16)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e16_fad = fad.score_clips(bg16, ev16)
    torch.cuda.synchronize()
    e16_s = wall_max((time.perf_counter() - t0) / e2e_steps)
    e2e_pcm16 = {"value": 2 * n_set / e16_s, "unit": "clips/s",
                 "h2d_bytes_per_step": int(bg16.numel() + ev16.numel()) * 2, "d2h_bytes_per_step": 8,
                 "api": "FrechetAudioDistance.score_clips(pinned int16 PCM)", "steps": e2e_steps, "fad": e2e16_fad}
    del bg16, ev16

    # ---- shard invariance on hardware: the same fixed 2 x 256 clips, (a) sharded over all ranks + NCCL all-reduce,
    # (b) scored by rank 0 alone.  SURVEY §4: must agree to 1e-10 (fp64 sums re-associate, nothing else changes).
    n_inv = 256
    inv_bg = gen_clips_gpu(0, 10_000_000, n_inv, dev)
    inv_ev = gen_clips_gpu(1, 10_000_000, n_inv, dev)
    ilo, ihi = shard_bounds(n_inv, rank, world)
    fad.process_group = None
    f_sharded = fad.score_clips(inv_bg[ilo:ihi], inv_ev[ilo:ihi])
    f_alone = fad.score_clips(inv_bg, inv_ev, reduce=False)
    shard_inv = abs(f_sharded - f_alone) / abs(f_alone)
    del inv_bg, inv_ev
    barrier()

    # ---- CPU baseline beside it, parity of every mode on the very same clips, the rest of the metric (rank 0, N = 1)
    cpu = frechet_ms = stats_ms = models = None
    if rank == 0 and world == 1:
        del bg, ev
        torch.cuda.empty_cache()
        if not args.no_cpu:
            r = cpu_oracle_rate(args.cpu_clips, args.cpu_clips)
            fad_cpu = r["fad"]
            from oracle import synth
            import numpy as np
            cb = torch.from_numpy(np.stack([synth.background_clip(i, CLIP_SAMPLES) for i in range(args.cpu_clips)]))
            ce = torch.from_numpy(np.stack([synth.eval_clip(i, CLIP_SAMPLES, 16000) for i in range(args.cpu_clips)]))
            per_mode = {}
            for prec in ("fp16x2", "bf16", "fp16", "bf16x3"):
                eng.set_precision(prec)
                f = fad.score_clips(cb, ce)
                per_mode[prec] = {"fad": f, "fad_rel_diff": abs(f - fad_cpu) / abs(fad_cpu)}
            eng.set_precision(args.precision)
            cpu = {"value": r["value"], "unit": "clips/s", "cores": r["threads"], "kind": "port",
                   "thread_config": r["config"], "calibration_clips_per_s": r["calibration_clips_per_s"],
                   "sample": f"{args.cpu_clips}+{args.cpu_clips} synthetic 10 s clips through oracle/pipeline.py "
                             f"(per-clip loop like fad.py:317, torch CPU fp32 + NumPy f64 front end), {r['s_per_step']:.1f} s",
                   "fad_cpu": fad_cpu, "fad_gpu_same_clips": per_mode[args.precision]["fad"],
                   "fad_rel_diff": per_mode[args.precision]["fad_rel_diff"], "fad_rel_diff_by_mode": per_mode}
        if not args.no_extras:
            frechet_ms, stats_ms = stats_frechet_microbench(eng, cpu=not args.no_cpu)
            del fad, eng
            torch.cuda.empty_cache()
            models = {m: model_throughput(m, args.model_clips, peaks, cpu=not args.no_cpu)
                      for m in ("pann-16k", "pann-32k", "clap")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE_NAMES[args.precision],
            "data": "synthetic (torch Philox on GPU, distributions of oracle/synth.py)",
            "config": workload_config(world, clips_per_set=args.clips_per_set),
            "fad": fad_value, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "frontend": frontend, "e2e_pcm16": e2e_pcm16, "cpu_baseline": cpu,
            "modes": modes, "shard_invariance_rel": shard_inv,
            "frechet_ms_d2048": frechet_ms["d2048"]["gpu_ms"] if frechet_ms else None,
            "frechet_ms": frechet_ms, "stats_ms": stats_ms, "models": models,
            "frac_of_tensor_roofline_whole_step": (value / world) * VGGISH_GFLOP_PER_CLIP / 1e3 / peaks["bf16_tflops"],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--cpu-worker":
        return _cpu_worker(json.loads(sys.argv[2]))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16x2", choices=["fp16x2", "fp16", "bf16", "bf16x3"])
    ap.add_argument("--clips-per-set", type=int, default=CLIPS_PER_SET_PER_GPU)
    ap.add_argument("--cpu-clips", type=int, default=64, help="clips per set of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-clips-ref", type=int, default=32, help="clips per set and step of the reference arm")
    ap.add_argument("--model-clips", type=int, default=512, help="ten-second clips per CNN14 model in `models`")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the Frechet / statistics microbench and `models`")
    ap.add_argument("--no-modes", action="store_true", help="skip the other precision modes")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
