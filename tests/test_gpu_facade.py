"""GPU tests of the drop-in façade: they read like the reference's own tests
(/root/reference/tests/test_basic.py, test_pann.py, test_clap.py) but run against
frechet_audio_distance_exported_b200.FrechetAudioDistance."""
import os

import numpy as np
import pytest
import torch

from oracle import networks, pipeline, stats, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fad_vgg():
    from frechet_audio_distance_exported_b200 import FrechetAudioDistance
    return FrechetAudioDistance(model_name="vggish", state_dict=networks.vggish_random_state_dict(seed=0))


def test_attributes_like_reference(fad_vgg):
    # fad.py:221-247
    assert fad_vgg.model_name == "vggish" and fad_vgg.sample_rate == 16000 and fad_vgg.channels == 1
    assert fad_vgg.verbose is False and fad_vgg.audio_load_worker == 8
    assert fad_vgg.device.type == "cuda" and os.path.isdir(fad_vgg.ckpt_dir)
    assert callable(fad_vgg.model)


def test_model_callable_contract(fad_vgg):
    # reference tests/test_basic.py:97-122: [B,1,96,64] -> [B,128] for several batch sizes
    for b in (1, 2, 10, 32):
        out = fad_vgg.model(torch.randn(b, 1, 96, 64))
        assert out.shape == (b, 128) and out.dtype == torch.float32
    with pytest.raises(ValueError):
        fad_vgg.model(torch.randn(2, 96, 64))


def test_frechet_distance_calculation(fad_vgg):
    # reference tests/test_basic.py:128-170
    d = fad_vgg.calculate_frechet_distance(np.array([1.0, 2.0, 3.0]), np.eye(3), np.array([1.0, 2.0, 3.0]), np.eye(3))
    assert abs(d) < 1e-6
    d = fad_vgg.calculate_frechet_distance(np.zeros(3), np.eye(3), np.ones(3), np.eye(3))
    assert d > 0
    with pytest.raises(AssertionError):                                    # fad.py:530-533
        fad_vgg.calculate_frechet_distance(np.zeros(3), np.eye(3), np.zeros(4), np.eye(4))


def test_embedding_statistics(fad_vgg):
    # reference tests/test_basic.py:172-190 (float64 input)
    np.random.seed(42)
    emb = np.random.randn(100, 128)
    mu, sigma = fad_vgg.calculate_embd_statistics(emb)
    assert mu.shape == (128,) and sigma.shape == (128, 128)
    np.testing.assert_allclose(mu, emb.mean(0), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(sigma, np.cov(emb, rowvar=False), rtol=1e-10, atol=1e-13)
    mu_l, _ = fad_vgg.calculate_embd_statistics([row for row in emb.astype(np.float32)])   # list input, fad.py:492-493
    assert mu_l.dtype == np.float32


def test_get_embeddings_semantics(fad_vgg):
    # fad.py:302-408: order preserved, mixed lengths, failed clips skipped, empty -> np.array([])
    a = synth.sine_clip(2.0, 440.0, 16000)
    b = synth.sine_clip(1.0, 880.0, 16000)
    short = synth.sine_clip(0.5, 440.0, 16000)
    stereo = np.stack([a, a], axis=1)
    out = fad_vgg.get_embeddings([a, b, short, stereo, a], 16000)
    assert out.shape == (2 + 1 + 0 + 2 + 2, 128)
    assert np.array_equal(out[0:2], out[5:7]) and np.array_equal(out[0:2], out[3:5])
    ora = pipeline.OracleFAD("vggish", networks.vggish_random_state_dict(seed=0))
    ref = ora.get_embeddings([a, b])
    assert np.max(np.abs(out[:3] - ref)) / np.max(np.abs(ref)) < 1e-2
    assert fad_vgg.get_embeddings([], 16000).shape == (0,)
    long44 = synth.sine_clip(3.0, 440.0, 44100)                            # 44.1 kHz input is resampled (vggish.py:249-250)
    e44 = fad_vgg.get_embeddings([long44], 44100)                          # resampled on the GPU (fadb_resample)
    assert e44.shape == (3, 128)
    from frechet_audio_distance_exported_b200.resample import resample
    assert np.array_equal(e44, fad_vgg.get_embeddings([resample(long44, 44100, 16000)], 16000))   # == host resampling
    assert fad_vgg.get_embeddings([long44[:10]], 44100).shape == (0, 128)  # too short for one patch after resampling
    assert fad_vgg._get_embedding_for_audio(a).shape == (2, 128)


def _write_wavs(d, clips, sr):
    from scipy.io import wavfile
    os.makedirs(d, exist_ok=True)
    for i, c in enumerate(clips):
        wavfile.write(os.path.join(d, f"{i:03d}.wav"), sr, np.round(c * 32767).astype(np.int16))
    open(os.path.join(d, ".hidden"), "w").close()                          # dot-files are skipped, fad.py:570


def test_score_directories_and_cache(tmp_path, fad_vgg):
    # reference scripts/verify_export.py:167-174 style: sine sets, 5 files each
    bg = [synth.sine_clip(2.0, 440.0 + 10 * i, 16000) + 0.01 * synth.background_clip(i, 32000) for i in range(5)]
    ev = [synth.sine_clip(2.0, 880.0 + 10 * i, 16000) + 0.01 * synth.background_clip(9 + i, 32000) for i in range(5)]
    _write_wavs(str(tmp_path / "bg"), bg, 16000)
    _write_wavs(str(tmp_path / "ev"), ev, 16000)
    cb, ce = str(tmp_path / "cache" / "bg.npy"), str(tmp_path / "cache" / "ev.npy")
    s = fad_vgg.score(str(tmp_path / "bg"), str(tmp_path / "ev"), background_embds_path=cb, eval_embds_path=ce)
    assert np.isfinite(s) and s > 0
    assert os.path.exists(cb) and np.load(cb).shape == (10, 128)
    s2 = fad_vgg.score("/nonexistent", "/nonexistent", background_embds_path=cb, eval_embds_path=ce)   # cache hit, fad.py:616-619
    assert s2 == s
    # the same files through the CPU oracle (PCM16-quantised like the WAVs)
    q = lambda c: (np.round(c * 32767).astype(np.int16).astype(np.float64) / 32768.0).astype(np.float32)
    ora = pipeline.OracleFAD("vggish", networks.vggish_random_state_dict(seed=0))
    names = sorted(f for f in os.listdir(str(tmp_path / "bg")) if not f.startswith("."))
    order = [int(f[:3]) for f in os.listdir(str(tmp_path / "bg")) if not f.startswith(".")]
    ref, _, _ = ora.fad_from_clips([q(bg[i]) for i in order], [q(ev[i]) for i in
                                   [int(f[:3]) for f in os.listdir(str(tmp_path / "ev")) if not f.startswith(".")]])
    assert len(names) == 5
    assert abs(s - ref) / abs(ref) < 2e-3                                  # default precision (fp16x2), tiny rank-deficient sets: 2x measured
    s16 = fad_vgg.score(str(tmp_path / "bg"), str(tmp_path / "ev"), dtype="int16")      # fad.py:145-149: raw PCM16 / 32768
    assert s16 == s                                                        # same samples, half the bytes over PCIe
    assert fad_vgg.score(str(tmp_path / "empty_missing"), str(tmp_path / "ev")) == -1     # exception -> -1, fad.py:660-662
    os.makedirs(str(tmp_path / "empty"))
    assert fad_vgg.score(str(tmp_path / "empty"), str(tmp_path / "ev")) == -1             # fad.py:640-642


@pytest.mark.parametrize("name,sr", [("pann-8k", 8000), ("pann-16k", 16000), ("pann-32k", 32000)])
def test_pann_preprocessing_to_model(name, sr):
    # reference tests/test_pann.py:161-197: (1, 2048) for 0.5 - 5 s clips
    from frechet_audio_distance_exported_b200 import FrechetAudioDistance
    sd = networks.cnn14_random_state_dict(seed=1)
    fad = FrechetAudioDistance(model_name=name, state_dict=sd, precision="bf16x3")
    ora = pipeline.OracleFAD(name, sd)
    clips = [synth.sine_clip(dur, 440.0, sr) + 0.05 * synth.background_clip(3, int(sr * dur)) for dur in (0.5, 1.0, 2.0)]
    out = fad.get_embeddings(clips, sr)
    assert out.shape == (3, 2048) and np.all(out >= 0)
    ref = ora.get_embeddings(clips)
    assert np.max(np.abs(out - ref)) / np.max(np.abs(ref)) < 1e-4
    out2 = fad.model(torch.randn(2, 1, 200, 64))                           # tests/test_pann.py:131-143 (any T)
    assert out2.shape == (2, 2048)


def test_clap_facade():
    # reference tests/test_clap.py:196-286 shape / L2 norm / determinism / FAD > 0 on 5+5 sine clips
    from frechet_audio_distance_exported_b200 import FrechetAudioDistance
    sd = networks.cnn14_random_state_dict(seed=2, clap_head=True)
    fad = FrechetAudioDistance(model_name="clap", state_dict=sd, precision="bf16x3")
    a = synth.sine_clip(2.0, 440.0, 48000)
    e1 = fad._get_embedding_for_audio(a)
    e2 = fad._get_embedding_for_audio(a)
    assert e1.shape == (1, 512) and np.array_equal(e1, e2)
    np.testing.assert_allclose(np.linalg.norm(e1, axis=1), 1.0, rtol=1e-5)
    ref = pipeline.OracleFAD("clap", sd).embed_clip(a)
    assert np.max(np.abs(e1 - ref)) / np.max(np.abs(ref)) < 1e-4
    bg = [synth.sine_clip(1.0, 440.0 + 10 * i, 48000) for i in range(5)]
    ev = [synth.sine_clip(1.0, 880.0 + 10 * i, 48000) for i in range(5)]
    mu1, s1 = fad.calculate_embd_statistics(fad.get_embeddings(bg, 48000))
    mu2, s2 = fad.calculate_embd_statistics(fad.get_embeddings(ev, 48000))
    f = fad.calculate_frechet_distance(mu1, s1, mu2, s2)
    assert np.isfinite(f) and f > 0
    # > 10 s: get_embeddings does not raise (only the reference's pad_audio_to_max_length helper does,
    # tests/test_clap.py:133-140); the log-mel is cut to 1001 frames, fad.py:87-89
    long = np.concatenate([a, a, a, a, a, a[:4801]]).astype(np.float32)
    assert long.shape[0] > 480000
    el = fad._get_embedding_for_audio(long)
    assert el.shape == (1, 512)
    assert np.max(np.abs(el - pipeline.OracleFAD("clap", sd).embed_clip(long))) / np.max(np.abs(ref)) < 1e-4


def test_ragged_clips_share_device_batches(fad_vgg):
    # clips whose lengths differ only past the last complete 0.96 s patch are one device batch (cut to the samples the
    # patches read); the embeddings are those of the clips embedded one by one
    rng = np.random.default_rng(3)
    lens = [16000 + 400 + int(k) for k in rng.integers(0, 15000, size=6)] + [2 * 16000 + 7, 3 * 16000 - 11]
    clips = [synth.background_clip(i, n) for i, n in enumerate(lens)]
    calls = []
    orig = fad_vgg._embed_group
    fad_vgg._embed_group = lambda c, rows, rs: (calls.append((len(c), c[0].shape[0])), orig(c, rows, rs))[1]
    try:
        out = fad_vgg.get_embeddings(clips, 16000)
    finally:
        fad_vgg._embed_group = orig
    assert len(calls) <= 3 and sum(c[0] for c in calls) == len(clips)          # 1-, 2- and 3-patch groups at most
    one = np.concatenate([fad_vgg.get_embeddings([c], 16000) for c in clips], axis=0)
    assert out.shape == one.shape and np.array_equal(out, one)
    ref = pipeline.OracleFAD("vggish", networks.vggish_random_state_dict(seed=0)).get_embeddings(clips)
    assert ref.shape == out.shape and np.max(np.abs(out - ref)) / np.max(np.abs(ref)) < 2.5e-3


def test_clap_foreign_sample_rate_follows_reference_order():
    # get_embeddings(x, sr != 48000): the reference pads to 480000 samples at the SOURCE rate (fad.py:355-359),
    # int16-truncates (clap.py:70-72), then resamples (clap.py:75-80) and keeps 1001 frames (fad.py:87-89)
    from frechet_audio_distance_exported_b200 import FrechetAudioDistance
    from frechet_audio_distance_exported_b200.resample import resample
    from oracle import frontend
    sd = networks.cnn14_random_state_dict(seed=2, clap_head=True)
    fad = FrechetAudioDistance(model_name="clap", state_dict=sd, precision="bf16x3")
    ora = pipeline.OracleFAD("clap", sd)
    sr = 16000
    clips = [(0.6 * synth.sine_clip(3.0, 440.0 + 60 * i, sr) + 0.05 * synth.background_clip(i, 3 * sr)).astype(np.float32)
             for i in range(2)]
    out = fad.get_embeddings(clips, sr)
    assert out.shape == (2, 512)
    for i, c in enumerate(clips):
        padded = np.pad(c, (0, 480000 - c.shape[0]))
        res = resample(frontend.clap_quantize(padded), sr, 48000).astype(np.float32)
        lm = frontend.pann_logmel(res, 48000)[:1001]
        ref = networks.clap_cnn14_forward(ora.sd, torch.from_numpy(lm)[None, None]).numpy()
        assert np.max(np.abs(out[i] - ref[0])) / np.max(np.abs(ref)) < 1e-4
    # the front end's own quantisation is back on afterwards: a native-rate clip still matches the oracle
    a = synth.sine_clip(1.0, 440.0, 48000)
    e = fad.get_embeddings([a], 48000)
    r = ora.embed_clip(a)
    assert np.max(np.abs(e - r)) / np.max(np.abs(r)) < 1e-4


def test_encodec_is_out_of_scope():
    from frechet_audio_distance_exported_b200 import FrechetAudioDistance
    with pytest.raises(NotImplementedError):
        FrechetAudioDistance(model_name="encodec-24k", state_dict={})


def test_ckpt_dir_with_exported_artefact(tmp_path):
    """FrechetAudioDistance(ckpt_dir) with a synthetic `vggish_exported.pt2` == the state_dict path."""
    from frechet_audio_distance_exported_b200 import FrechetAudioDistance
    from test_abi import _vggish_module
    sd = networks.vggish_random_state_dict(seed=3)
    ep = torch.export.export(_vggish_module(sd), (torch.randn(2, 1, 96, 64),))
    torch.export.save(ep, str(tmp_path / "vggish_exported.pt2"))
    a = FrechetAudioDistance(str(tmp_path), "vggish")                     # positional order of the reference, fad.py:178-186
    b = FrechetAudioDistance(model_name="vggish", state_dict=sd)
    clip = synth.eval_clip(0, 32400, 16000)
    assert np.array_equal(a._get_embedding_for_audio(clip), b._get_embedding_for_audio(clip))


def test_missing_weights_fail_loudly(tmp_path):
    from frechet_audio_distance_exported_b200 import FrechetAudioDistance
    with pytest.raises(FileNotFoundError):
        FrechetAudioDistance(ckpt_dir=str(tmp_path), model_name="vggish")
    from frechet_audio_distance_exported_b200._lib import FadbError
    with pytest.raises(FadbError, match="missing weight tensor"):
        FrechetAudioDistance(model_name="vggish", state_dict={"features.0.weight": torch.zeros(64, 1, 3, 3)})
