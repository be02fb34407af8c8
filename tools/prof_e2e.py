"""Profiling helper (not a test): end-to-end score_clips from pinned host PCM for a few streaming chunk sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from frechet_audio_distance_exported_b200 import FrechetAudioDistance
from oracle import networks

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6250
fad = FrechetAudioDistance(model_name="vggish", state_dict=networks.vggish_random_state_dict(0))
eng, d = fad.engine, 128
g = torch.Generator().manual_seed(1)
bg = (torch.randn(n, 160000, generator=g) * 0.1).clamp_(-1, 1).pin_memory()
ev = (torch.randn(n, 160000, generator=g) * 0.2).clamp_(-1, 1).pin_memory()


def score(chunk):
    both = torch.zeros(2 * (1 + d + d * d), dtype=torch.float64, device=fad.device)
    half = both.numel() // 2
    fad.accumulate_clips(bg, both[:half], chunk_clips=chunk)
    fad.accumulate_clips(ev, both[half:], chunk_clips=chunk)
    mu1, s1 = eng.stats_finalize(both[:half], d)
    mu2, s2 = eng.stats_finalize(both[half:], d)
    return float(eng.frechet(mu1, s1, mu2, s2)[0].item())


for chunk in (1024, 768, 512, 384, 256):
    score(chunk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        f = score(chunk)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 2
    print(f"chunk {chunk}: {dt * 1e3:.1f} ms -> {2 * n / dt:.0f} clips/s (fad {f:.4f})", flush=True)
