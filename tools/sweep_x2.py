"""Sensitivity sweep (not a test): which tensor-core layers need the lo weight plane (fp16x2) for FAD parity?
For a set of FADB_X2_MASK values: FAD deviation against the CPU oracle on 64 + 64 ten-second clips (several weight
seeds) and the time of the tensor-core layers.  usage: sweep_x2.py [n_clips] [seeds]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from frechet_audio_distance_exported_b200 import FrechetAudioDistance
from oracle import networks, pipeline, synth

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
seeds = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
torch.set_num_threads(len(os.sched_getaffinity(0)))
N = 160000
bg = np.stack([synth.background_clip(i, N) for i in range(n_clips)])
ev = np.stack([synth.eval_clip(i, N, 16000) for i in range(n_clips)])
tb, te = torch.from_numpy(bg).cuda(), torch.from_numpy(ev).cuda()
big = torch.cat([tb, te] * 8)                       # 16 x n_clips clips for timing
masks = [("all", 0xff), ("none", 0x00), ("convs", 0x1f), ("fcs", 0xe0)] + [(f"only{i}", 1 << i) for i in range(8)] + \
        [(f"without{i}", 0xff ^ (1 << i)) for i in range(8)]
for seed in seeds:
    sd = networks.vggish_random_state_dict(seed=seed)
    t0 = time.time()
    ref, _, _ = pipeline.OracleFAD("vggish", sd).fad_from_clips(list(bg), list(ev))
    print(f"seed {seed}: oracle FAD {ref:.6f} ({time.time() - t0:.1f} s)", flush=True)
    for prec in ("bf16", "fp16", "bf16x3"):
        fad = FrechetAudioDistance(model_name="vggish", state_dict=sd, precision=prec)
        f = fad.score_clips(tb, te)
        print(json.dumps({"seed": seed, "precision": prec, "fad_rel_diff": abs(f - ref) / ref}), flush=True)
        del fad
    for name, m in masks:
        os.environ["FADB_X2_MASK"] = hex(m)
        fad = FrechetAudioDistance(model_name="vggish", state_dict=sd, precision="fp16x2")
        f = fad.score_clips(tb, te)
        eng = fad.engine
        eng.embed_pcm(big); torch.cuda.synchronize()
        eng.profile_enable(True)
        eng.embed_pcm(big); torch.cuda.synchronize()
        ms, fl, n = eng.profile_read()
        eng.profile_enable(False)
        print(json.dumps({"seed": seed, "mask": name, "fad_rel_diff": abs(f - ref) / ref, "signed": (f - ref) / ref,
                          "gemm_ms_per_1k_clips": ms / (big.shape[0] / 1000.0)}), flush=True)
        del fad, eng
    os.environ.pop("FADB_X2_MASK", None)
