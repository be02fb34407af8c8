"""Profiling helper (not a test): VGGish pcm->emb throughput for overlap on/off and GEMM smem budgets.
Each configuration runs in its own process (env vars are read at handle creation)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from frechet_audio_distance_exported_b200.engine import Engine
    from oracle import networks
    eng = Engine("vggish", networks.vggish_random_state_dict(seed=0), precision="bf16")
    pcm = (torch.randn(2040, 160000, device="cuda") * 0.1).clamp(-1, 1)
    ref = None
    for _ in range(2):
        out = eng.embed_pcm(pcm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        out = eng.embed_pcm(pcm)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"overlap={os.environ.get('FADB_OVERLAP','1')} smem={os.environ.get('FADB_GEMM_SMEM','196608')} lowprio={os.environ.get('FADB_AUX_PRIO','1')}: "
          f"{ms:.2f} ms / 2040 clips -> {2040 / ms * 1e3:.0f} clips/s  checksum {out.double().sum().item():.6f}")
else:
    for ov, sm, pr in (("0", "196608", "1"), ("1", "196608", "0"), ("1", "196608", "1"), ("1", "147456", "1"), ("1", "147456", "0")):
        env = dict(os.environ, FADB_OVERLAP=ov, FADB_GEMM_SMEM=sm, FADB_AUX_PRIO=pr)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=600)
        print(r.stdout.strip() or r.stderr[-1500:])
