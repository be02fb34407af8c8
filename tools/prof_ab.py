"""Profiling helper (not a test): same-box A/B of env-var switches for VGGish pcm->emb throughput.
usage: python tests/prof_ab.py VAR=a,b [VAR2=c,d ...]   (each combination runs in its own process)"""
import itertools, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from frechet_audio_distance_exported_b200.engine import Engine
    from oracle import networks
    mb = int(os.environ.get("AB_MAX_BATCH", "2048"))
    eng = Engine("vggish", networks.vggish_random_state_dict(seed=0), precision="bf16", max_batch=mb)
    g = torch.Generator(device="cuda").manual_seed(1)
    pcm = (torch.randn(2040, 160000, device="cuda", generator=g) * 0.1).clamp(-1, 1)
    for _ in range(2):
        out = eng.embed_pcm(pcm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(4):
        out = eng.embed_pcm(pcm)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print(f"{sys.argv[2]}: {ms:.2f} ms / 2040 clips -> {2040 / ms * 1e3:.0f} clips/s  checksum {out.double().sum().item():.4f}")
else:
    axes = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[1:]]
    for combo in itertools.product(*[v for _, v in axes]):
        env = dict(os.environ)
        tag = " ".join(f"{k}={v}" for (k, _), v in zip(axes, combo))
        for (k, _), v in zip(axes, combo):
            env[k] = v
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", tag], env=env, capture_output=True, text=True, timeout=900)
        print(r.stdout.strip() or r.stderr[-1500:], flush=True)
