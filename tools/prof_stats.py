"""Profiling helper (not a test): one statistics accumulation at d = 2048 (tensor-core syrk path), for an ncu launch list.
usage: prof_stats.py [rows] [d]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from frechet_audio_distance_exported_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
d = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
eng = Engine("vggish")
eng.set_tensor_syrk(True)
x = torch.randn(n, d, device="cuda") + 0.5
for _ in range(2):
    acc = eng.new_acc(d)
    eng.stats_accumulate(x, acc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
acc = eng.new_acc(d)
eng.stats_accumulate(x, acc)
e1.record(); torch.cuda.synchronize()
print(f"stats_accumulate {n} x {d}: {e0.elapsed_time(e1):.3f} ms")
