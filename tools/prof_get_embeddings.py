"""Profiling helper (not a test): throughput of the reference-shaped call `get_embeddings(list of numpy clips, sr)` on
one GPU — host gather into pinned staging, ring copies, embedding, results back — for equal-length and ragged lists.
usage: prof_get_embeddings.py [clips]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from frechet_audio_distance_exported_b200 import FrechetAudioDistance
from oracle import networks

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
fad = FrechetAudioDistance(model_name="vggish", state_dict=networks.vggish_random_state_dict(0))
rng = np.random.default_rng(0)
base = (rng.standard_normal(160000 + 4096) * 0.1).astype(np.float32)
equal = [np.ascontiguousarray(base[i % 4096: i % 4096 + 160000]) for i in range(n)]
ragged = [np.ascontiguousarray(base[: 16400 + int(k)]) for k in rng.integers(0, 143000, size=n)]
for name, clips in (("equal 10 s", equal), ("ragged 1-10 s", ragged)):
    fad.get_embeddings(clips[:64], 16000)
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        e = fad.get_embeddings(clips, 16000)
        dt = time.perf_counter() - t
        secs = sum(c.shape[0] for c in clips) / 16000
        print(f"{name}: {n} clips in {dt * 1e3:.0f} ms = {n / dt:.0f} clips/s, {secs / dt:.0f} audio-seconds/s, "
              f"embeddings {e.shape}", flush=True)
