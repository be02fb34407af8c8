"""Profiling helper (not a test): BASELINE configs[4], the Frechet-statistics microbench — mean/cov accumulation and the
fp64 Frechet kernel chain at d = 128 / 512 / 2048, N = 1e4 .. 1e6 embeddings, next to numpy / scipy on the host cores
(bounded: np.cov up to 1e5 rows, scipy sqrtm once per d).  Embeddings: x = z A + b as in SURVEY 8d.
usage: bench_stats.py [--no-cpu]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from frechet_audio_distance_exported_b200.engine import Engine
from oracle import networks, stats

cpu = "--no-cpu" not in sys.argv
eng = Engine("vggish")
eng.set_tensor_syrk(True)                                # the opt-in tensor-core syrk (d >= 512, >= 8192 rows)
eng64 = Engine("vggish")                                 # the default fp64 DFMA syrk: the checker of the tensor-core path
dev = eng.device


def make(n, d, seed, scale, shift):
    g = torch.Generator(device=dev).manual_seed(seed)
    a = torch.randn(d, d, generator=g, device=dev) / d ** 0.5
    out = torch.empty(n, d, device=dev)
    for i in range(0, n, 65536):                         # bounded temporaries
        m = min(65536, n - i)
        out[i:i + m] = (torch.randn(m, d, generator=g, device=dev) @ a) * scale + shift
    return out


def gpu_ms(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


for d in (128, 512, 2048):
    cpu_sqrtm_s = None
    for n in (10_000, 100_000, 1_000_000):
        x1, x2 = make(n, d, 1, 1.0, 0.0), make(n, d, 2, 1.1, 0.05)

        def stats_both():
            out = []
            for x in (x1, x2):
                acc = eng.new_acc(d)
                eng.stats_accumulate(x, acc)
                out.append(eng.stats_finalize(acc, d))
            return out
        ms_stats, ((mu1, s1), (mu2, s2)) = gpu_ms(stats_both)
        ms_fr, fr = gpu_ms(lambda: eng.frechet(mu1, s1, mu2, s2))

        def stats_both64():
            out = []
            for x in (x1, x2):
                acc = eng64.new_acc(d)
                eng64.stats_accumulate(x, acc)
                out.append(eng64.stats_finalize(acc, d))
            return out
        ms64, ((m1_, c1_), (m2_, c2_)) = gpu_ms(stats_both64, reps=1)
        fr64 = eng64.frechet(m1_, c1_, m2_, c2_)
        tc_vs_64 = {"fp64_kernel_stats_ms_both_sets": ms64,
                    "sigma_rel_err_vs_fp64_kernel": float((s1 - c1_).abs().max() / c1_.abs().max()),
                    "fad_rel_err_vs_fp64_kernel": abs(float(fr[0]) - float(fr64[0])) / abs(float(fr64[0]))}
        rec = {"d": d, "n_per_set": n, "gpu_stats_ms_both_sets": ms_stats, "gpu_frechet_ms": ms_fr, "fad": float(fr[0]),
               "stats_rows_per_s": 2 * n / ms_stats * 1e3, "stats_read_gbs": 2 * n * d * 4 / ms_stats / 1e6,
               "stats_tflops": 2 * n * d * (d + 1) / ms_stats / 1e9, **tc_vs_64}
        if cpu and n <= 100_000:
            h1, h2 = x1.cpu().numpy(), x2.cpu().numpy()
            t0 = time.perf_counter()
            m1, c1 = stats.embd_statistics(h1)
            m2, c2 = stats.embd_statistics(h2)
            rec["cpu_numpy_stats_s"] = time.perf_counter() - t0
            rec["sigma_max_rel_err"] = float(np.abs(s1.cpu().numpy() - c1).max() / np.abs(c1).max())
            if cpu_sqrtm_s is None or d <= 512:
                t0 = time.perf_counter()
                ref = stats.frechet_distance(m1.astype(np.float64), c1, m2.astype(np.float64), c2)
                cpu_sqrtm_s = time.perf_counter() - t0
                rec["cpu_scipy_frechet_s"] = cpu_sqrtm_s
                rec["fad_rel_err_vs_cpu"] = abs(float(fr[0]) - ref) / abs(ref)
        print(json.dumps(rec), flush=True)
        del x1, x2
        torch.cuda.empty_cache()
