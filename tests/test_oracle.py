"""CPU tests: the oracle restatement against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py), the reference's own known-answer tests, and an independent second opinion
(torch.stft + torchaudio Slaney filterbank) for the un-vendored librosa pieces."""
import numpy as np
import pytest
import torch

from oracle import frontend, networks, pipeline, stats, synth


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)


def test_vggish_frontend_matches_reference(golden):
    z = golden("vggish_frontend.npz")
    clips = {
        "sine440_1s": synth.sine_clip(1.0, 440.0, 16000),
        "sine880_2s": synth.sine_clip(2.0, 880.0, 16000),
        "bg7_2p5s": synth.background_clip(7, 40000),
        "ev3_2p5s": synth.eval_clip(3, 40000, 16000),
        "short_0p5s": synth.sine_clip(0.5, 440.0, 16000),
    }
    for k, c in clips.items():
        out = frontend.vggish_examples(c)
        assert out.shape == z[k].shape, k
        assert out.dtype == np.float32
        if out.size:
            assert relerr(out, z[k]) < 1e-6, k


def test_vggish_shapes_like_reference_tests():
    # reference tests/test_basic.py:30-66
    assert frontend.vggish_examples(synth.sine_clip(2.0, 440.0, 16000)).shape == (2, 96, 64)
    assert frontend.vggish_examples(synth.sine_clip(0.5, 440.0, 16000)).shape == (0, 96, 64)
    for n, p in ((8000, 0), (15600, 1), (16000, 1), (31120, 2), (160000, 10)):
        assert frontend.vggish_num_patches(n) == p
    stereo = np.stack([synth.sine_clip(1.0, 440.0, 16000)] * 2, axis=1)
    np.testing.assert_allclose(frontend.vggish_examples(stereo), frontend.vggish_examples(stereo[:, 0]), atol=1e-6)


def test_vggish_core_matches_reference(golden):
    z = golden("vggish_core.npz")
    sd = networks.vggish_random_state_dict(seed=int(z["seed"]))
    out = networks.vggish_forward(sd, torch.from_numpy(z["patches"])[:, None]).numpy()
    assert out.shape == (5, 128)
    assert relerr(out, z["embeddings"]) < 1e-5


def test_vggish_e2e_matches_reference(golden):
    z = golden("vggish_e2e.npz")
    n, k = int(z["n_samples"]), int(z["n_clips"])
    sd = networks.vggish_random_state_dict(seed=0)
    of = pipeline.OracleFAD("vggish", sd)
    bg = [synth.background_clip(i, n) for i in range(k)]
    ev = [synth.eval_clip(i, n, 16000) for i in range(k)]
    fad, eb, ee = of.fad_from_clips(bg, ev)
    assert relerr(eb, z["emb_bg"]) < 1e-5 and relerr(ee, z["emb_ev"]) < 1e-5
    assert abs(fad - float(z["fad"])) / float(z["fad"]) < 1e-6


def test_pann_clap_frontend_matches_reference(golden):
    z = golden("pann_frontend.npz")
    for sr in (8000, 16000, 32000):
        out = frontend.pann_features(synth.eval_clip(11, sr, sr), sr)
        assert out.shape == z[f"pann_{sr}"].shape == (104, 64)       # T = 101 -> 104 = 32*4 - 24
        assert relerr(out, z[f"pann_{sr}"]) < 1e-6
        assert np.all(out[101:] == 0.0)
    out = frontend.clap_features(synth.eval_clip(12, 48000, 48000))
    assert out.shape == (1001, 64)
    assert relerr(out, z["clap_48000"]) < 1e-6


def test_pann_frame_counts():
    for sr, hop in ((8000, 80), (16000, 160), (32000, 320)):
        t = frontend.pann_num_frames(10 * sr, hop)
        assert t == 1001 and frontend.pann_padded_frames(t) == 1032
    assert [frontend.pann_padded_frames(t) for t in (8, 9, 40, 41, 1001)] == [8, 40, 40, 72, 1032]


def test_librosa_restatement_against_torch_second_opinion():
    """librosa is un-vendored (parity unpinned): cross-check the restated STFT / Slaney filterbank
    against torch.stft and torchaudio."""
    torchaudio = pytest.importorskip("torchaudio")
    for sr, cfg in frontend.PANN_CONFIGS.items():
        fb = frontend.slaney_mel_filterbank(sr, cfg["n_fft"], 64, cfg["fmin"], cfg["fmax"])
        ref = torchaudio.functional.melscale_fbanks(cfg["n_fft"] // 2 + 1, cfg["fmin"], cfg["fmax"], 64, sr,
                                                    norm="slaney", mel_scale="slaney").T.numpy()
        assert np.max(np.abs(fb - ref)) < 1e-5 * np.max(np.abs(ref))   # torchaudio builds it in float32
        x = synth.eval_clip(5, sr // 2, sr)
        p = frontend.stft_power_centered(x, cfg["n_fft"], cfg["hop"])
        st = torch.stft(torch.from_numpy(x).double(), cfg["n_fft"], cfg["hop"], cfg["n_fft"],
                        window=torch.hann_window(cfg["n_fft"], periodic=True, dtype=torch.float64),
                        center=True, pad_mode="reflect", return_complex=True)
        pr = (st.abs() ** 2).numpy()
        assert p.shape == pr.shape
        assert relerr(p, pr) < 1e-5


def test_clap_quantisation_truncates_toward_zero():
    x = np.array([0.99999, -0.99999, 0.5, -0.5, 1.0 / 32767 * 0.9], dtype=np.float32)
    q = frontend.clap_quantize(x) * 32767
    np.testing.assert_allclose(q, [32766, -32766, 16383, -16383, 0], atol=1e-3)


def test_cnn14_core_matches_reference(golden):
    z = golden("cnn14_core.npz")
    sd = networks.cnn14_random_state_dict(seed=int(z["seed"]))
    out = networks.cnn14_forward(sd, torch.from_numpy(z["feats"])[:, None]).numpy()
    assert out.shape == (2, 2048) and np.all(out >= 0)
    assert relerr(out, z["embeddings"]) < 1e-5


def test_clap_head_is_l2_normalised():
    sd = networks.cnn14_random_state_dict(seed=2, clap_head=True)
    x = torch.randn(1, 1, 104, 64) * 10 - 30
    out = networks.clap_cnn14_forward(sd, x).numpy()
    assert out.shape == (1, 512)
    np.testing.assert_allclose(np.linalg.norm(out, axis=1), 1.0, rtol=1e-5)       # reference tests/test_clap.py:225-240


def test_stats_and_frechet_match_reference(golden):
    z = golden("stats_frechet.npz")
    for tag in ("d16", "d128", "d64_singular"):
        n, d = z[f"{tag}_shape"]
        a, b = synth.embedding_set(0, int(n), int(d)), synth.embedding_set(1, int(n), int(d))
        mu1, s1 = stats.embd_statistics(a)
        mu2, s2 = stats.embd_statistics(b)
        assert mu1.dtype == np.float32 and s1.dtype == np.float64
        np.testing.assert_allclose(s1, z[f"{tag}_sigma1"], rtol=1e-12, atol=1e-14)
        fd = stats.frechet_distance(mu1, s1, mu2, s2)
        assert abs(fd - float(z[f"{tag}_fd"])) / float(z[f"{tag}_fd"]) < 1e-9
        assert abs(stats.frechet_distance_eigh(mu1, s1, mu2, s2) - fd) / fd < 1e-6


def test_reference_known_answer_tests():
    # reference tests/test_basic.py:143-170
    assert abs(stats.frechet_distance(np.array([1., 2., 3.]), np.eye(3), np.array([1., 2., 3.]), np.eye(3))) < 1e-6
    assert stats.frechet_distance(np.zeros(3), np.eye(3), np.ones(3), np.eye(3)) > 0
    np.random.seed(42)
    mu, sigma = stats.embd_statistics(np.random.randn(100, 128))                   # test_basic.py:183-190
    assert mu.shape == (128,) and sigma.shape == (128, 128)
    with pytest.raises(AssertionError):
        stats.frechet_distance(np.zeros(3), np.eye(3), np.zeros(4), np.eye(4))


def test_oracle_skips_bad_clips_like_reference():
    sd = networks.vggish_random_state_dict(seed=0)
    of = pipeline.OracleFAD("vggish", sd)
    assert of.get_embeddings([]).shape == (0,)                                     # fad.py:405-406
    out = of.get_embeddings([synth.sine_clip(0.5, 440.0, 16000), synth.sine_clip(1.0, 440.0, 16000)])
    assert out.shape == (1, 128)


def test_config_size_goldens_match_oracle(golden):
    """tests/golden/vggish_e2e_10s.npz, cnn14_10s.npz (oracle/make_golden.py --configs: the UNMODIFIED reference on
    ten-second clips): the oracle restatement reproduces them (bounded: 2 VGGish clips, 1 PANN-16k clip, Frechet at d = 128)."""
    z = golden("vggish_e2e_10s.npz")
    sd = networks.vggish_random_state_dict(seed=0)
    ora = pipeline.OracleFAD("vggish", sd)
    n = int(z["n_samples"])
    emb = ora.get_embeddings([synth.background_clip(i, n) for i in range(2)])
    assert relerr(emb, z["emb_bg_head"][:20]) < 1e-5
    fd = stats.frechet_distance(z["mu1"], z["sigma1"], z["mu2"], z["sigma2"])
    assert abs(fd - float(z["fad"])) / float(z["fad"]) < 1e-9
    c = golden("cnn14_10s.npz")
    sdp = networks.cnn14_random_state_dict(seed=int(c["seed"]), clap_head=True)
    e = pipeline.OracleFAD("pann-16k", sdp).get_embeddings([synth.eval_clip(21, 160000, 16000)])
    assert e.shape == (1, 2048) and relerr(e, c["pann_16k"]) < 1e-5


def test_host_ring_chunking():
    from frechet_audio_distance_exported_b200.stream import chunk_bounds
    assert chunk_bounds(10, 4) == [(0, 4), (4, 4), (8, 2)]
    assert chunk_bounds(10, 4, first=2) == [(0, 2), (2, 4), (6, 4)]
    assert chunk_bounds(0, 4) == [] and chunk_bounds(3, 8) == [(0, 3)]
    b = chunk_bounds(12500, 512, first=256)
    assert sum(n for _, n in b) == 12500 and all(b[i][0] + b[i][1] == b[i + 1][0] for i in range(len(b) - 1))


def test_threaded_gather_into_staging_rows():
    # fad.py::_gather_clips: the four-thread copy of a clip list into the rows of a staging buffer is the plain loop
    from frechet_audio_distance_exported_b200.fad import _gather_clips
    rng = np.random.default_rng(0)
    for n, length, dt in [(3, 1000, np.float32), (64, 4096, np.float32), (37, 777, np.int16)]:
        clips = [(rng.standard_normal(length) * 1000).astype(dt) for _ in range(n + 5)]
        view = np.zeros((n, length), dtype=dt)
        _gather_clips(view, clips, 2, n)
        assert np.array_equal(view, np.stack(clips[2:2 + n]))
