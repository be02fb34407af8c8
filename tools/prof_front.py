"""Profiling helper (not a test): duration of the fused VGGish front-end + conv1 kernel via the profile hook.
usage: prof_front.py [clips]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from frechet_audio_distance_exported_b200.engine import Engine
from oracle import networks

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1020
eng = Engine("vggish", networks.vggish_random_state_dict(0), max_batch=16384)
g = torch.Generator(device="cuda").manual_seed(1)
pcm = (torch.randn(n, 160000, device="cuda", generator=g) * 0.1).clamp(-1, 1)
for _ in range(3):
    eng.embed_pcm(pcm)
torch.cuda.synchronize()
eng.profile_enable(True)
for _ in range(3):
    eng.embed_pcm(pcm)
torch.cuda.synchronize()
ms, fl, nl = eng.profile_read()
fm = eng.front_ms / 3
gb = n * (640000 + 10 * 48 * 32 * 64 * 2) / 1e9
print(f"DBG={os.environ.get('FADB_FRONT_DBG', '0')}: front+conv1 {fm:.3f} ms / {n} clips "
      f"({n / fm * 1e3:.0f} clips/s, {gb / fm * 1e3:.0f} GB/s of PCM-in + bf16-out); tensor layers {ms / 3:.3f} ms ({fl / ms / 1e9:.0f} TFLOP/s)")
