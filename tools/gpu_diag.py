"""GPU bring-up diagnostics: every kernel family against the CPU oracle, one section per process so a
faulting kernel cannot take the others down.  Prints a table instead of asserting.

    python tests/gpu_diag.py            # all sections, each in its own subprocess with a timeout
    python tests/gpu_diag.py conv       # one section in-process
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SECTIONS = ["conv", "frontend", "vggish", "stats", "frechet", "cnn14", "perf"]


def relerr(a, b):
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def sec_conv():
    import numpy as np
    import torch
    import torch.nn.functional as F
    from frechet_audio_distance_exported_b200.engine import Engine

    eng = Engine("vggish")
    cases = [
        # name, B, H, W, Cin, Cout, k, relu, pool
        ("linear 200x128->64", 1, 1, 200, 128, 64, 1, 0, 0),
        ("linear 300x256->256", 1, 1, 300, 256, 256, 1, 1, 0),
        ("conv 2x8x16 64->128", 2, 8, 16, 64, 128, 3, 1, 0),
        ("conv 2x8x16 64->128 max", 2, 8, 16, 64, 128, 3, 1, 1),
        ("conv 2x8x16 64->128 avg", 2, 8, 16, 64, 128, 3, 1, 2),
        ("conv 3x24x16 128->256", 3, 24, 16, 128, 256, 3, 1, 0),
        ("conv 5x12x8 512->512 max", 5, 12, 8, 512, 512, 3, 1, 1),
        ("conv 2x48x32 64->128 max", 2, 48, 32, 64, 128, 3, 1, 1),
        ("conv 1x129x8 64->128 avg", 1, 129, 8, 64, 128, 3, 1, 2),
        ("conv 2x32x2 128->256", 2, 32, 2, 128, 256, 3, 1, 0),
        ("conv 1x40x64 64->64 avg", 1, 40, 64, 64, 64, 3, 1, 2),
        ("halo 2x32x16 128->256", 2, 32, 16, 128, 256, 3, 1, 0),
        ("halo 1x64x8 192->64 max", 1, 64, 8, 192, 64, 3, 1, 1),
        ("halo 3x60x16 64->128 avg", 3, 60, 16, 64, 128, 3, 1, 2),
    ]
    g = torch.Generator().manual_seed(1)
    for prec in ("bf16", "bf16x3"):
        eng.set_precision(prec)
        for name, B, H, W, Cin, Cout, k, relu, pool in cases:
            x = torch.randn((B, H, W, Cin), generator=g)
            w = torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5
            b = torch.randn(Cout, generator=g) * 0.1
            if prec == "bf16":
                xr, wr = x.bfloat16().float(), w.bfloat16().float()
            else:
                xr, wr = x, w
            ref = F.conv2d(xr.permute(0, 3, 1, 2).double(), wr.double(), b.double(), padding=k // 2)
            if relu:
                ref = F.relu(ref)
            if pool == 1:
                ref = F.max_pool2d(ref, 2)
            elif pool == 2:
                ref = F.avg_pool2d(ref, 2)
            ref = ref.permute(0, 2, 3, 1).numpy()
            try:
                out = eng.debug_conv_layer(x.cuda(), w.cuda(), b.cuda(), k, bool(relu), pool).cpu().numpy()
                print(f"conv[{prec:6s}] {name:28s} rel-max-err {relerr(out, ref):.3e}  shape {out.shape}")
            except Exception as e:
                print(f"conv[{prec:6s}] {name:28s} FAILED: {e}")
                if "device-side" in str(e) or "CUDA" in str(e) or "cuda" in str(e):
                    return


def sec_frontend():
    import numpy as np
    import torch
    from frechet_audio_distance_exported_b200.engine import Engine
    from oracle import frontend, synth

    z = np.load(os.path.join(ROOT, "tests", "golden", "vggish_frontend.npz"))
    eng = Engine("vggish")
    clips = {"sine440_1s": synth.sine_clip(1.0, 440.0, 16000), "sine880_2s": synth.sine_clip(2.0, 880.0, 16000),
             "bg7_2p5s": synth.background_clip(7, 40000), "ev3_2p5s": synth.eval_clip(3, 40000, 16000)}
    for k, c in clips.items():
        out = eng.frontend(torch.from_numpy(c)[None].cuda()).cpu().numpy()
        print(f"frontend vggish {k:12s} vs golden rel-max-err {relerr(out, z[k]):.3e} shape {out.shape}")
    zp = np.load(os.path.join(ROOT, "tests", "golden", "pann_frontend.npz"))
    for name, sr in (("pann-8k", 8000), ("pann-16k", 16000), ("pann-32k", 32000)):
        e = Engine(name)
        c = synth.eval_clip(11, sr, sr)
        out = e.frontend(torch.from_numpy(c)[None].cuda()).cpu().numpy()[0]
        print(f"frontend {name:9s} vs golden rel-max-err {relerr(out, zp[f'pann_{sr}']):.3e} "
              f"max-abs dB {np.abs(out - zp[f'pann_{sr}']).max():.3e} shape {out.shape}")
        s = synth.sine_clip(1.0, 440.0, sr)
        out = e.frontend(torch.from_numpy(s)[None].cuda()).cpu().numpy()[0]
        ref = frontend.pann_features(s, sr)
        print(f"frontend {name:9s} sine     rel-max-err {relerr(out, ref):.3e} max-abs dB {np.abs(out - ref).max():.3e}")
    e = Engine("clap")
    c = synth.eval_clip(12, 48000, 48000)
    out = e.frontend(torch.from_numpy(c)[None].cuda()).cpu().numpy()[0]
    print(f"frontend clap      vs golden rel-max-err {relerr(out, zp['clap_48000']):.3e} "
          f"max-abs dB {np.abs(out - zp['clap_48000']).max():.3e} shape {out.shape}")


def sec_vggish():
    import numpy as np
    import torch
    from frechet_audio_distance_exported_b200.engine import Engine
    from oracle import networks

    z = np.load(os.path.join(ROOT, "tests", "golden", "vggish_core.npz"))
    sd = networks.vggish_random_state_dict(seed=0)
    for prec in ("bf16", "bf16x3"):
        eng = Engine("vggish", sd, precision=prec)
        out = eng.embed_features(torch.from_numpy(z["patches"]).cuda()).cpu().numpy()
        print(f"vggish core [{prec:6s}] embeddings vs golden rel-max-err {relerr(out, z['embeddings']):.3e}")
        e2 = np.load(os.path.join(ROOT, "tests", "golden", "vggish_e2e.npz"))
        from oracle import synth
        n = int(e2["n_samples"])
        bg = np.stack([synth.background_clip(i, n) for i in range(4)])
        out = eng.embed_pcm(torch.from_numpy(bg).cuda()).cpu().numpy()
        print(f"vggish e2e  [{prec:6s}] pcm->emb   vs golden rel-max-err {relerr(out, e2['emb_bg']):.3e}")


def sec_stats():
    import numpy as np
    import torch
    from frechet_audio_distance_exported_b200.engine import Engine
    from oracle import synth

    eng = Engine("vggish")
    for n, d in ((1000, 128), (5000, 512), (300, 200), (257, 2048)):
        x = synth.embedding_set(0, n, d)
        acc = eng.new_acc(d)
        t = torch.from_numpy(x).cuda()
        eng.stats_accumulate(t[: n // 3], acc)
        eng.stats_accumulate(t[n // 3:], acc)
        mu, sg = eng.stats_finalize(acc, d)
        print(f"stats n={n} d={d}: mu err {relerr(mu.cpu().numpy(), x.astype(np.float64).mean(0)):.3e} "
              f"sigma err {relerr(sg.cpu().numpy(), np.cov(x, rowvar=False)):.3e}")


def sec_frechet():
    import numpy as np
    import torch
    from frechet_audio_distance_exported_b200.engine import Engine
    from oracle import stats, synth

    eng = Engine("vggish")
    z = np.load(os.path.join(ROOT, "tests", "golden", "stats_frechet.npz"))
    for tag in ("d16", "d128", "d64_singular"):
        t = [torch.from_numpy(np.ascontiguousarray(z[f"{tag}_{k}"], dtype=np.float64)).cuda()
             for k in ("mu1", "sigma1", "mu2", "sigma2")]
        out = eng.frechet(*t).cpu().numpy()
        print(f"frechet {tag:13s} {out[0]:.10f} vs golden {float(z[tag + '_fd']):.10f} rel {abs(out[0] - z[tag + '_fd']) / abs(z[tag + '_fd']):.3e}")
    for mu_a, mu_b in ((np.array([1., 2., 3.]), np.array([1., 2., 3.])), (np.zeros(3), np.ones(3))):
        t = [torch.from_numpy(a).cuda() for a in (mu_a, np.eye(3), mu_b, np.eye(3))]
        print("frechet KAT", eng.frechet(*t).cpu().numpy())
    for n, d in ((2000, 512), (1000, 2048), (10000, 2048)):
        a = synth.embedding_set(0, n, d)
        b = synth.embedding_set(1, n, d)
        m1, s1 = stats.embd_statistics(a)
        m2, s2 = stats.embd_statistics(b)
        t = [torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).cuda() for v in (m1, s1, m2, s2)]
        eng.frechet(*t)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = eng.frechet(*t).cpu().numpy()
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        ref = stats.frechet_distance_eigh(m1, s1, m2, s2)
        dt_ref = time.perf_counter() - t1
        print(f"frechet n={n} d={d}: gpu {out[0]:.8f} ({dt * 1e3:.1f} ms) vs cpu-eigh {ref:.8f} ({dt_ref * 1e3:.0f} ms) rel {abs(out[0] - ref) / abs(ref):.3e}")


def sec_cnn14():
    import numpy as np
    import torch
    from frechet_audio_distance_exported_b200.engine import Engine
    from oracle import networks

    z = np.load(os.path.join(ROOT, "tests", "golden", "cnn14_core.npz"))
    sd = networks.cnn14_random_state_dict(seed=1)
    for prec in ("bf16", "bf16x3"):
        eng = Engine("pann-16k", sd, precision=prec)
        out = eng.embed_features(torch.from_numpy(z["feats"]).cuda()).cpu().numpy()
        print(f"cnn14 core [{prec:6s}] embeddings vs golden rel-max-err {relerr(out, z['embeddings']):.3e}")
    sdc = networks.cnn14_random_state_dict(seed=2, clap_head=True)
    eng = Engine("clap", sdc, precision="bf16x3")
    x = torch.randn(2, 1001, 64, generator=torch.Generator().manual_seed(3)) * 10 - 30
    out = eng.embed_features(x.cuda()).cpu().numpy()
    ref = networks.clap_cnn14_forward(sdc, x[:, None]).numpy()
    print(f"clap cnn14+head [bf16x3] vs oracle rel-max-err {relerr(out, ref):.3e} norms {np.linalg.norm(out, axis=1)}")


def sec_perf():
    import numpy as np
    import torch
    from frechet_audio_distance_exported_b200.engine import Engine
    from oracle import networks

    sd = networks.vggish_random_state_dict(seed=0)
    eng = Engine("vggish", sd, precision="bf16", max_batch=2048)

    def timeit(fn, n=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    pcm = (torch.randn(200, 160000, device="cuda") * 0.1).clamp(-1, 1)
    ms = timeit(lambda: eng.frontend(pcm))
    print(f"perf frontend vggish 200 clips: {ms:.3f} ms  -> {200 / ms * 1e3:.0f} clips/s, {200 * 885760 / ms / 1e6:.1f} GB/s")
    feats = eng.frontend(pcm)
    for prec in ("bf16", "bf16x3"):
        eng.set_precision(prec)
        ms = timeit(lambda: eng.embed_features(feats))
        print(f"perf vggish core [{prec}] 2000 patches: {ms:.3f} ms -> {200 / ms * 1e3:.0f} clips/s, "
              f"{2000 * 1.7278e9 / ms / 1e9:.1f} TFLOP/s")
    eng.set_precision("bf16")
    ms = timeit(lambda: eng.embed_pcm(pcm))
    print(f"perf vggish pcm->emb 200 clips: {ms:.3f} ms -> {200 / ms * 1e3:.0f} clips/s")


def main():
    if len(sys.argv) > 1:
        globals()["sec_" + sys.argv[1]]()
        return
    for s in SECTIONS:
        print(f"===== {s} =====", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), s], timeout=600, capture_output=True, text=True)
            print(r.stdout, end="")
            if r.returncode != 0:
                print(f"[section {s} exit {r.returncode}]\n" + r.stderr[-3000:])
        except subprocess.TimeoutExpired as e:
            print(f"[section {s} TIMED OUT]", (e.stdout or b"")[-2000:] if e.stdout else "")
        print(f"({time.time() - t0:.1f} s)", flush=True)


if __name__ == "__main__":
    main()
