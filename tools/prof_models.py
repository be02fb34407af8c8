"""Profiling helper (not a test): throughput of every model family of BASELINE.json configs on one GPU
(device-resident PCM, CUDA events): PCM -> embeddings -> stats -> Frechet, clips/s and tensor TFLOP/s."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from frechet_audio_distance_exported_b200.engine import Engine
from oracle import networks

GFLOP = {"vggish": 17.278, "pann-8k": 41.3627, "pann-16k": 41.3627, "pann-32k": 41.3627, "clap": 40.0843}
SR = {"vggish": 16000, "pann-8k": 8000, "pann-16k": 16000, "pann-32k": 32000, "clap": 48000}
models = sys.argv[1].split(",") if len(sys.argv) > 1 else ["pann-16k", "pann-32k", "clap", "vggish"]
n_clips = int(sys.argv[2]) if len(sys.argv) > 2 else 256
out = []
for m in models:
    sd = networks.vggish_random_state_dict(0) if m == "vggish" else networks.cnn14_random_state_dict(1, clap_head=(m == "clap"))
    for prec in ("bf16",):
        eng = Engine(m, sd, precision=prec, max_batch=(2048 if m == "vggish" else 64))
        g = torch.Generator(device="cuda").manual_seed(1)
        pcm = (torch.randn(n_clips, 10 * SR[m], device="cuda", generator=g) * 0.1).clamp(-1, 1)
        d = eng.dim

        def step():
            acc = eng.new_acc()
            eng.stats_accumulate(eng.embed_pcm(pcm), acc)
            mu, sg = eng.stats_finalize(acc, d)
            return eng.frechet(mu, sg, mu, sg)

        step(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(2):
            r = step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        # embed only
        e0.record(); eng.embed_pcm(pcm); e1.record(); torch.cuda.synchronize()
        ms_e = e0.elapsed_time(e1)
        rec = {"model": m, "precision": prec, "clips": n_clips, "ms_embed_stats_frechet": ms, "ms_embed": ms_e,
               "clips_per_s_embed": n_clips / ms_e * 1e3, "tensor_tflops_embed": n_clips * GFLOP[m] / ms_e,
               "fad_self": float(r[0])}
        print(json.dumps(rec), flush=True)
        del eng, pcm
        torch.cuda.empty_cache()
