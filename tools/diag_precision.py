"""Diagnostic (not a test): embedding error of every precision mode against an fp64 torch forward on the GPU, for several
batch sizes and kernel-path switches (pairs / halo / resident weights).  usage: diag_precision.py [vggish|cnn14]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 2 and sys.argv[2] == "child":
    import numpy as np, torch
    from frechet_audio_distance_exported_b200.engine import Engine
    from oracle import networks
    which = sys.argv[1]
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if which == "vggish":
        sd = networks.vggish_random_state_dict(0)
        sizes = (2, 18, 40, 300)
        make = lambda n: torch.randn(n, 96, 64, generator=torch.Generator().manual_seed(n)).cuda() * 2.0
        fwd = lambda sdd, x: networks.vggish_forward(sdd, x[:, None])
        name = "vggish"
    else:
        sd = networks.cnn14_random_state_dict(1)
        sizes = (1, 3, 40)
        make = lambda n: (torch.randn(n, 200, 64, generator=torch.Generator().manual_seed(n)) * 10 - 30).cuda()
        fwd = lambda sdd, x: networks.cnn14_forward(sdd, x[:, None])
        name = "pann-16k"
    sdd = {k: (v.double().cuda() if v.is_floating_point() else v.cuda()) for k, v in sd.items()}
    out = []
    for prec in ("bf16", "fp16", "fp16x2", "bf16x3"):
        eng = Engine(name, sd, precision=prec)
        row = []
        for n in sizes:
            x = make(n)
            ref = fwd(sdd, x.double()).cpu().numpy()
            e = eng.embed_features(x).cpu().numpy()
            err = np.abs(e - ref).max() / np.abs(ref).max()
            bias = float(((e - ref) * np.sign(ref)).mean() / np.abs(ref).mean())
            row.append(f"n={n}: {err:.2e} (shrink {bias:+.1e})")
        out.append(f"  {prec:7s} " + " | ".join(row))
        del eng
    print("\n".join(out), flush=True)
    sys.exit(0)

which = sys.argv[1] if len(sys.argv) > 1 else "vggish"
for env in ({}, {"FADB_CLUSTER": "0"}, {"FADB_CLUSTER": "2"}, {"FADB_HALO": "0"}, {"FADB_RESIDENT_B": "0"}):
    e = dict(os.environ); e.update(env)
    print(f"== {which} {env}", flush=True)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), which, "child"], env=e, capture_output=True, text=True)
    print(r.stdout + (r.stderr[-1500:] if r.returncode else ""), flush=True)
