"""Generate tests/golden/*.npz by running the UNMODIFIED reference (through oracle/ref_shim.py).

Run in the builder container only (needs /root/reference):

    python -m oracle.make_golden

Each fixture stores the seeded inputs' identity (so tests regenerate the inputs with
oracle/synth.py), the reference's outputs, and — as a self-check executed here — the oracle
restatement is compared against the reference before anything is written.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import frontend, networks, pipeline, ref_shim, stats, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _check(name, a, b, tol):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30) if a.size else 0.0
    status = "ok" if err <= tol else "MISMATCH"
    print(f"  oracle vs reference  {name:40s} rel-max-err {err:.3e}  (tol {tol:g})  {status}")
    if err > tol:
        raise SystemExit(f"oracle restatement disagrees with the reference on {name}")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = ref_shim.load_reference()
    from frechet_audio_distance_exported.models import vggish as rvgg, pann as rpann, clap as rclap
    from frechet_audio_distance_exported import fad as rfad

    os.makedirs(OUT, exist_ok=True)

    # ---------------------------------------------------------------- VGGish front end
    clips = {
        "sine440_1s": synth.sine_clip(1.0, 440.0, 16000),
        "sine880_2s": synth.sine_clip(2.0, 880.0, 16000),
        "bg7_2p5s": synth.background_clip(7, 40000),
        "ev3_2p5s": synth.eval_clip(3, 40000, 16000),
        "short_0p5s": synth.sine_clip(0.5, 440.0, 16000),
    }
    fe = {}
    for k, c in clips.items():
        r = rvgg.waveform_to_examples(c, 16000, return_tensor=True).numpy()[:, 0]
        o = frontend.vggish_examples(c)
        assert r.shape == o.shape, (k, r.shape, o.shape)
        _check(f"vggish_frontend/{k}", o, r, 1e-6)
        fe[k] = r.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "vggish_frontend.npz"), **fe)

    # ---------------------------------------------------------------- VGGishCore
    sd = networks.vggish_random_state_dict(seed=0)
    model = rvgg.VGGishCore()
    model.load_state_dict(sd)
    model.eval()
    patches = np.concatenate([fe["bg7_2p5s"], fe["ev3_2p5s"], fe["sine440_1s"]], axis=0)   # 5 patches
    x = torch.from_numpy(patches)[:, None]
    with torch.no_grad():
        r_emb = model(x).numpy()
        r_feat = model.features(x).numpy()
    o_emb, acts = networks.vggish_forward(sd, x, return_intermediates=True)
    _check("vggish_core/embeddings", o_emb.numpy(), r_emb, 1e-5)
    np.savez_compressed(
        os.path.join(OUT, "vggish_core.npz"), patches=patches, embeddings=r_emb,
        features_mean=r_feat.mean(axis=(0, 2, 3)), seed=np.int64(0))

    # ---------------------------------------------------------------- VGGish end to end (reference get_embeddings + FAD)
    rf = ref_shim.make_reference_fad(ref, "vggish", model)
    n = 3 * 16000 + 400
    bg = [synth.background_clip(i, n) for i in range(4)]
    ev = [synth.eval_clip(i, n, 16000) for i in range(4)]
    e_bg = rf.get_embeddings(bg, 16000)
    e_ev = rf.get_embeddings(ev, 16000)
    mu1, s1 = rf.calculate_embd_statistics(e_bg)
    mu2, s2 = rf.calculate_embd_statistics(e_ev)
    fad_ref = float(rf.calculate_frechet_distance(mu1, s1, mu2, s2))
    of = pipeline.OracleFAD("vggish", sd)
    fad_o, o_bg, o_ev = of.fad_from_clips(bg, ev)
    _check("vggish_e2e/emb_bg", o_bg, e_bg, 1e-5)
    _check("vggish_e2e/emb_ev", o_ev, e_ev, 1e-5)
    _check("vggish_e2e/fad", fad_o, fad_ref, 1e-6)
    np.savez_compressed(os.path.join(OUT, "vggish_e2e.npz"), n_samples=np.int64(n), n_clips=np.int64(4),
                        emb_bg=e_bg, emb_ev=e_ev, fad=np.float64(fad_ref))

    # ---------------------------------------------------------------- PANN / CLAP front ends (librosa stubbed: PARITY UNPINNED)
    pf = {}
    for sr in (8000, 16000, 32000):
        c = synth.eval_clip(11, sr, sr)            # 1 s
        r = rpann.waveform_to_logmel(c, sr, target_sample_rate=sr, return_tensor=True)
        r = rfad._pad_to_valid_pann_time(r).numpy()[0, 0]
        o = frontend.pann_features(c, sr)
        assert r.shape == o.shape, (sr, r.shape, o.shape)
        _check(f"pann_frontend/{sr}", o, r, 1e-6)
        pf[f"pann_{sr}"] = r
    c = synth.eval_clip(12, 48000, 48000)
    cp = np.pad(c, (0, 480000 - c.shape[0]))
    r = rclap.preprocess_for_clap(cp, 48000, return_tensor=True)
    r = rfad._pad_to_clap_time(r).numpy()[0, 0]
    o = frontend.clap_features(c)
    _check("clap_frontend/48000", o, r, 1e-6)
    pf["clap_48000"] = r
    np.savez_compressed(os.path.join(OUT, "pann_frontend.npz"), **pf)

    # ---------------------------------------------------------------- PANNCore (CNN14)
    sdp = networks.cnn14_random_state_dict(seed=1)
    pm = rpann.PANNCore()
    pm.load_state_dict(sdp)
    pm.eval()
    xs = torch.from_numpy(np.stack([pf["pann_16000"], pf["pann_32000"]]))[:, None]    # [2,1,104,64]
    with torch.no_grad():
        r_emb = pm(xs).numpy()
    o_emb = networks.cnn14_forward(sdp, xs).numpy()
    _check("cnn14_core/embeddings", o_emb, r_emb, 1e-5)
    np.savez_compressed(os.path.join(OUT, "cnn14_core.npz"), feats=xs.numpy()[:, 0], embeddings=r_emb,
                        seed=np.int64(1))

    # ---------------------------------------------------------------- statistics + Frechet
    st = {}
    for tag, (n_rows, d) in {"d16": (200, 16), "d128": (1000, 128), "d64_singular": (40, 64)}.items():
        a = synth.embedding_set(0, n_rows, d)
        b = synth.embedding_set(1, n_rows, d)
        mu1, s1 = rf.calculate_embd_statistics(a)
        mu2, s2 = rf.calculate_embd_statistics(b)
        fd = float(rf.calculate_frechet_distance(mu1, s1, mu2, s2))
        omu1, os1 = stats.embd_statistics(a)
        omu2, os2 = stats.embd_statistics(b)
        _check(f"stats/{tag}/sigma", os1, s1, 1e-12)
        _check(f"stats/{tag}/fd", stats.frechet_distance(omu1, os1, omu2, os2), fd, 1e-9)
        _check(f"stats/{tag}/fd_eigh", stats.frechet_distance_eigh(omu1, os1, omu2, os2), fd, 1e-6)
        st[f"{tag}_mu1"], st[f"{tag}_sigma1"], st[f"{tag}_mu2"], st[f"{tag}_sigma2"] = mu1, s1, mu2, s2
        st[f"{tag}_fd"] = np.float64(fd)
        st[f"{tag}_shape"] = np.array([n_rows, d], dtype=np.int64)
    # reference known-answer tests (tests/test_basic.py:143-170 of the reference)
    kat0 = float(rf.calculate_frechet_distance(np.array([1., 2., 3.]), np.eye(3), np.array([1., 2., 3.]), np.eye(3)))
    kat1 = float(rf.calculate_frechet_distance(np.zeros(3), np.eye(3), np.ones(3), np.eye(3)))
    assert abs(kat0) < 1e-6 and kat1 > 0
    st["kat_identical"], st["kat_shifted"] = np.float64(kat0), np.float64(kat1)
    np.savez_compressed(os.path.join(OUT, "stats_frechet.npz"), **st)
    print("golden fixtures written to", OUT)


def main_configs():
    """BASELINE.json configs as quoted (ten-second clips): the UNMODIFIED reference's get_embeddings / statistics /
    Frechet on seeded clips.  Writes vggish_e2e_10s.npz, cnn14_10s.npz (separate from main() so the small fixtures
    are not regenerated):

        python -m oracle.make_golden --configs
    """
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = ref_shim.load_reference()
    from frechet_audio_distance_exported.models import vggish as rvgg, pann as rpann, clap as rclap
    from frechet_audio_distance_exported import fad as rfad
    os.makedirs(OUT, exist_ok=True)

    # ---- VGGish, 48 + 48 ten-second clips (BASELINE configs[0] shape): FAD and statistics of the reference
    sd = networks.vggish_random_state_dict(seed=0)
    model = rvgg.VGGishCore()
    model.load_state_dict(sd)
    model.eval()
    rf = ref_shim.make_reference_fad(ref, "vggish", model)
    n, k = 160000, 48
    bg = [synth.background_clip(i, n) for i in range(k)]
    ev = [synth.eval_clip(i, n, 16000) for i in range(k)]
    e_bg, e_ev = rf.get_embeddings(bg, 16000), rf.get_embeddings(ev, 16000)
    mu1, s1 = rf.calculate_embd_statistics(e_bg)
    mu2, s2 = rf.calculate_embd_statistics(e_ev)
    fad_ref = float(rf.calculate_frechet_distance(mu1, s1, mu2, s2))
    of = pipeline.OracleFAD("vggish", sd)
    fad_o, o_bg, o_ev = of.fad_from_clips(bg[:4], ev[:4])
    _check("vggish_e2e_10s/emb_bg[:4]", o_bg, e_bg[:40], 1e-5)
    np.savez_compressed(os.path.join(OUT, "vggish_e2e_10s.npz"), n_samples=np.int64(n), n_clips=np.int64(k),
                        emb_bg_head=e_bg[:40], emb_ev_head=e_ev[:40], mu1=mu1, sigma1=s1, mu2=mu2, sigma2=s2,
                        fad=np.float64(fad_ref))
    print(f"  vggish 10 s: {k}+{k} clips, FAD {fad_ref:.6f}")

    # ---- CNN14 models, one ten-second clip each (T = 1001 -> 1032 / 1001) + PANN-16k 12 + 12 clips FAD
    out = {}
    sdp = networks.cnn14_random_state_dict(seed=1, clap_head=True)
    pm = rpann.PANNCore()
    pm.load_state_dict({k_: v for k_, v in sdp.items() if not k_.startswith("clap_head")})
    pm.eval()
    for name, sr in (("pann-8k", 8000), ("pann-16k", 16000), ("pann-32k", 32000)):
        rfp = ref_shim.make_reference_fad(ref, name, pm)
        c = synth.eval_clip(21, 10 * sr, sr)
        e = rfp.get_embeddings([c], sr)                                       # fad.py:372-385 incl. the time padding
        o = pipeline.OracleFAD(name, sdp).get_embeddings([c])
        _check(f"cnn14_10s/{name}", o, e, 1e-5)
        out[name.replace("-", "_")] = e
    c = synth.eval_clip(22, 480000, 48000)
    x = rfad._pad_to_clap_time(rclap.preprocess_for_clap(c, 48000, return_tensor=True))   # fad.py:356-362
    with torch.no_grad():
        h = pm(x)                                                             # the reference's PANNCore
        h = torch.nn.functional.relu(torch.nn.functional.linear(h, sdp["clap_head.0.weight"], sdp["clap_head.0.bias"]))
        h = torch.nn.functional.linear(h, sdp["clap_head.2.weight"], sdp["clap_head.2.bias"])
        e = torch.nn.functional.normalize(h, dim=-1).numpy()                  # README.md:195-199 (no reference code)
    _check("cnn14_10s/clap", pipeline.OracleFAD("clap", sdp).get_embeddings([c]), e, 1e-5)
    out["clap"] = e
    rfp = ref_shim.make_reference_fad(ref, "pann-16k", pm)
    k = 12
    bg = [synth.background_clip(100 + i, 160000) for i in range(k)]
    ev = [synth.eval_clip(100 + i, 160000, 16000) for i in range(k)]
    e_bg, e_ev = rfp.get_embeddings(bg, 16000), rfp.get_embeddings(ev, 16000)
    mu1, s1 = rfp.calculate_embd_statistics(e_bg)
    mu2, s2 = rfp.calculate_embd_statistics(e_ev)
    fad_ref = float(rfp.calculate_frechet_distance(mu1, s1, mu2, s2))
    out.update(pann16k_emb_bg=e_bg, pann16k_emb_ev=e_ev, pann16k_fad=np.float64(fad_ref), pann16k_clips=np.int64(k),
               seed=np.int64(1))
    np.savez_compressed(os.path.join(OUT, "cnn14_10s.npz"), **out)
    print(f"  pann-16k 10 s: {k}+{k} clips, FAD {fad_ref:.6f}")
    print("config fixtures written to", OUT)


if __name__ == "__main__":
    sys.exit(main_configs() if "--configs" in sys.argv else main())
