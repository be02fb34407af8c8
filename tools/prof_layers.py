"""Profiling helper (not a test): per-layer time of the tensor-core layers via fadb_debug_conv_layer-sized
launches driven through the engine's profile hook.  usage: prof_layers.py [vggish|cnn14]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from frechet_audio_distance_exported_b200.engine import Engine
from oracle import networks

which = sys.argv[1] if len(sys.argv) > 1 else "vggish"
if which == "vggish":
    eng = Engine("vggish", networks.vggish_random_state_dict(0), max_batch=4096)
    x = torch.randn(4000, 96, 64, device="cuda")
else:
    eng = Engine("pann-16k", networks.cnn14_random_state_dict(1), max_batch=64)
    x = torch.randn(64, 1032, 64, device="cuda") * 10 - 30
for _ in range(2):
    eng.embed_features(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(3):
    eng.embed_features(x)
e1.record(); torch.cuda.synchronize()
eng.profile_enable(True)
eng.embed_features(x)
torch.cuda.synchronize()
ms, fl, n = eng.profile_read()
print(f"{which} skipA={os.environ.get('FADB_DEBUG_SKIP_A','0')} resB={os.environ.get('FADB_RESIDENT_B','1')}: "
      f"network {e0.elapsed_time(e1) / 3:.3f} ms; tensor-core layers {ms:.3f} ms ({n} launches, {fl / ms / 1e9:.0f} TFLOP/s)")
