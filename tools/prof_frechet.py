"""Profiling helper (not a test): one Frechet call at d = 2048 (for `ncu` launch lists), plus wall time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from frechet_audio_distance_exported_b200.engine import Engine
from oracle import stats, synth

d = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
eng = Engine("vggish")
a, b = synth.embedding_set(0, n, d), synth.embedding_set(1, n, d)
ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
acc1, acc2 = eng.new_acc(d), eng.new_acc(d)
eng.stats_accumulate(ta, acc1); eng.stats_accumulate(tb, acc2)
m1, s1 = eng.stats_finalize(acc1, d); m2, s2 = eng.stats_finalize(acc2, d)
torch.cuda.synchronize()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); out = eng.frechet(m1, s1, m2, s2); e1.record(); torch.cuda.synchronize()
    print(f"frechet d={d} n={n}: {out[0].item():.8f}  {e0.elapsed_time(e1):.2f} ms")
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); acc1.zero_(); eng.stats_accumulate(ta, acc1); e1.record(); torch.cuda.synchronize()
print(f"stats accumulate n={n} d={d}: {e0.elapsed_time(e1):.2f} ms")
