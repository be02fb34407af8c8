"""Profiling helper (not a test): tensor-core layer time of every precision mode on the SAME box and in one process
(boxes differ by a few per cent): bf16, fp16, fp16x2 with the e4m3 low-order pass, fp16x2 with the fp16 low-order pass.
usage: prof_modes.py [clips] [vggish|pann-16k]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from frechet_audio_distance_exported_b200.engine import Engine
from oracle import networks

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1638
model = sys.argv[2] if len(sys.argv) > 2 else "vggish"
sd = networks.vggish_random_state_dict(0) if model == "vggish" else networks.cnn14_random_state_dict(1)
sr = 16000
g = torch.Generator(device="cuda").manual_seed(1)
pcm = (torch.randn(n, 10 * sr, device="cuda", generator=g) * 0.1).clamp(-1, 1)
for rep in range(2):
    for name, prec, env in (("bf16", "bf16", {}), ("fp16", "fp16", {}), ("fp16x2 (e4m3 lo pass)", "fp16x2", {"FADB_LO_FP8": "1"}),
                            ("fp16x2 (e4m3, not Cin=64)", "fp16x2", {"FADB_LO_FP8": "1", "FADB_LO_FP8_C64": "0"}),
                            ("fp16x2 (fp16 lo pass)", "fp16x2", {"FADB_LO_FP8": "0"})):
        os.environ.update(env)
        eng = Engine(model, sd, precision=prec)
        for _ in range(2):
            eng.embed_pcm(pcm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            eng.embed_pcm(pcm)
        e1.record(); torch.cuda.synchronize()
        eng.profile_enable(True)
        eng.embed_pcm(pcm); torch.cuda.synchronize()
        ms, fl, nl = eng.profile_read()
        eng.profile_enable(False)
        print(f"[{rep}] {model} {name:24s}: embed {e0.elapsed_time(e1) / 3:8.3f} ms / {n} clips; tensor layers {ms:8.3f} ms "
              f"({fl / ms / 1e9:6.0f} algorithmic TFLOP/s), front {eng.front_ms:.3f} ms", flush=True)
        del eng
        for k in env:
            os.environ.pop(k, None)
