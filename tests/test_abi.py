"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/fadb.h declares; host-only entry points behave; nothing in the product imports the oracle;
the product fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fadb.h")
PKG = os.path.join(ROOT, "frechet_audio_distance_exported_b200")


@pytest.fixture(scope="module")
def lib():
    from frechet_audio_distance_exported_b200 import _lib
    return _lib.load()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fadb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from frechet_audio_distance_exported_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    bound = {n for n, _, _ in _lib.SYMBOLS}
    assert set(names) == bound, set(names) ^ bound            # the ctypes table mirrors the header exactly
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fadb_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert lib.fadb_abi_version() == 1


def test_library_is_sm100a_tcgen05_tma():
    """the built .so carries sm_100a SASS with tcgen05 MMA, TMEM loads and TMA loads"""
    from frechet_audio_distance_exported_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnem in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnem in sass, mnem
    assert "UTCHMMA.2CTA" in sass                      # the CTA-pair variant (cta_group::2) is built in
    assert "STG.E.ENL2.256" in sass                    # 32-byte epilogue stores


def test_frontend_rows_matches_oracle(lib):
    from oracle import frontend
    for n in (0, 399, 400, 8000, 15600, 16000, 31120, 40000, 160000):
        assert lib.fadb_frontend_rows(0, n) == frontend.vggish_num_patches(n)
    for model, sr, hop in ((1, 8000, 80), (2, 16000, 160), (3, 32000, 320)):
        for n in (sr // 2, sr, 10 * sr, 10 * sr + 7):
            assert lib.fadb_frontend_rows(model, n) == frontend.pann_padded_frames(frontend.pann_num_frames(n, hop))
    assert lib.fadb_frontend_rows(4, 480000) == 1001 and lib.fadb_frontend_rows(4, 1000) == 1001
    assert lib.fadb_frontend_rows(9, 1000) == -1
    assert [lib.fadb_embed_dim(m) for m in range(5)] == [128, 2048, 2048, 2048, 512]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    h = C.c_void_p()
    rc = lib.fadb_create(C.byref(h), 0)
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.fadb_last_error() or b"CUDA" in lib.fadb_last_error()
    from frechet_audio_distance_exported_b200 import Engine, FrechetAudioDistance
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine("vggish")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FrechetAudioDistance(model_name="vggish", state_dict={})


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+\.*oracle\b|importlib.*oracle|/oracle/", re.M)
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not pat.search(txt), f"{f} references the oracle"


def test_constructor_validation_like_reference():
    # fad.py:205-219 — raised before any device work, so checkable on CPU
    from frechet_audio_distance_exported_b200 import FrechetAudioDistance, VALID_MODELS
    with pytest.raises(ValueError, match="Unknown model"):
        FrechetAudioDistance(model_name="nope")
    with pytest.raises(ValueError, match="requires sample_rate=16000"):
        FrechetAudioDistance(model_name="vggish", sample_rate=44100)
    assert VALID_MODELS["clap"] == {"sample_rate": 48000, "embedding_dim": 512}      # reference tests/test_clap.py:170-176
    assert set(VALID_MODELS) == {"vggish", "pann-8k", "pann-16k", "pann-32k", "encodec-24k", "encodec-48k", "clap"}


def test_padding_helpers_like_reference():
    from frechet_audio_distance_exported_b200 import CLAP_TIME_FRAMES, _pad_to_clap_time, _pad_to_valid_pann_time
    assert CLAP_TIME_FRAMES == 1001
    assert _pad_to_valid_pann_time(torch.ones(1, 1, 1001, 64)).shape[2] == 1032          # fad.py:41-66
    x = _pad_to_valid_pann_time(torch.ones(1, 1, 41, 64))
    assert x.shape[2] == 72 and float(x[0, 0, 41:].abs().sum()) == 0.0
    assert _pad_to_clap_time(torch.ones(1, 1, 900, 64)).shape[2] == 1001                 # fad.py:69-91
    assert _pad_to_clap_time(torch.ones(1, 1, 1200, 64)).shape[2] == 1001


def test_load_audio_wav(tmp_path):
    # reference tests/test_basic.py:196-247 (PCM16 round trip to 4 decimals, stereo -> mono)
    from scipy.io import wavfile
    from frechet_audio_distance_exported_b200 import load_audio
    from oracle import synth
    a = synth.sine_clip(1.0, 440.0, 16000)
    p = str(tmp_path / "a.wav")
    wavfile.write(p, 16000, np.round(a * 32767).astype(np.int16))
    out = load_audio(p, 16000, 1)
    assert out.dtype == np.float32 and len(out) == len(a)
    np.testing.assert_array_almost_equal(out, a, decimal=4)
    p2 = str(tmp_path / "s.wav")
    wavfile.write(p2, 16000, np.stack([np.round(a * 32767).astype(np.int16)] * 2, axis=1))
    out2 = load_audio(p2, 16000, 1)
    assert out2.ndim == 1
    np.testing.assert_array_almost_equal(out2, a, decimal=4)
    out3 = load_audio(p, 16000, 1, dtype="int16")                                       # fad.py:148-149
    assert out3.dtype == np.float64 and np.abs(out3 - a).max() < 1e-4
    raw = load_audio(p, 16000, 1, dtype="int16", raw_pcm16=True)                        # B200 extension: raw PCM16
    assert raw.dtype == np.int16 and np.array_equal(raw / 32768.0, out3)
    assert load_audio(p2, 16000, 1, dtype="int16", raw_pcm16=True).dtype == np.float64  # stereo: host path
    assert load_audio(p, 8000, 1, dtype="int16", raw_pcm16=True).dtype == np.float64    # needs resampling: host path
    # resampling on load (reference tests/test_basic.py:212-228: 1 s at 44.1 kHz -> exactly 16000 samples)
    t = np.linspace(0, 1.0, 44100, dtype=np.float32)
    b = (np.sin(2 * np.pi * 440.0 * t) * 0.5).astype(np.float32)
    p3 = str(tmp_path / "f.wav")
    wavfile.write(p3, 44100, b)
    out4 = load_audio(p3, 16000, 1)
    assert len(out4) == 16000


def test_resampler_properties():
    """resampy semantics (un-vendored, parity unpinned): output length int(n * ratio), identity at equal rates,
    a band-limited tone survives up- and down-sampling."""
    from frechet_audio_distance_exported_b200.resample import resample
    x = np.sin(2 * np.pi * 440.0 * np.arange(16000) / 16000)
    assert np.array_equal(resample(x, 16000, 16000), x)
    up = resample(x, 16000, 48000)
    assert len(up) == 48000
    ref = np.sin(2 * np.pi * 440.0 * np.arange(48000) / 48000)
    assert np.abs(up[6000:42000] - ref[6000:42000]).max() < 1e-6
    x44 = np.sin(2 * np.pi * 440.0 * np.arange(44100) / 44100).astype(np.float32)
    down = resample(x44, 44100, 16000)
    assert len(down) == 16000 and down.dtype == np.float32
    ref = np.sin(2 * np.pi * 440.0 * np.arange(16000) / 16000)
    assert np.abs(down[2000:14000] - ref[2000:14000]).max() < 5e-3      # resampy's truncated table step: 0.27 % gain
    hi = np.sin(2 * np.pi * 15000.0 * np.arange(44100) / 44100)          # above the new Nyquist: must be removed
    assert np.abs(resample(hi, 44100, 16000)[2000:14000]).max() < 1e-3
    assert len(resample(np.zeros(7), 8000, 16000)) == 14
    with pytest.raises(ValueError):
        resample(np.zeros(1), 48000, 8000)


def _vggish_module(sd):
    """nn.Module with VGGishCore's parameter names (features.N / embeddings.N), built from the oracle's layer list."""
    import torch.nn as nn
    from oracle.networks import VGGISH_CFG
    layers, cin = [], 1
    for v in VGGISH_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(2, 2))
        else:
            layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU()]
            cin = v

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.features = nn.Sequential(*layers)
            self.embeddings = nn.Sequential(nn.Linear(12288, 4096), nn.ReLU(), nn.Linear(4096, 4096), nn.ReLU(),
                                            nn.Linear(4096, 128))

        def forward(self, x):
            x = self.features(x).permute(0, 2, 3, 1).reshape(x.shape[0], -1)
            return self.embeddings(x)

    m = M()
    m.load_state_dict(sd)
    return m.eval()


def test_weights_from_exported_pt2(tmp_path):
    """the reference loads `vggish_exported.pt2` (fad.py:297); we read the same artefact for its
    state_dict.  A synthetic artefact is produced here exactly like scripts/export_vggish.py does
    (torch.export.export + save) because the real one cannot be downloaded offline."""
    from frechet_audio_distance_exported_b200.fad import FrechetAudioDistance
    from oracle import networks
    sd = networks.vggish_random_state_dict(seed=3)
    ep = torch.export.export(_vggish_module(sd), (torch.randn(2, 1, 96, 64),))
    torch.export.save(ep, str(tmp_path / "vggish_exported.pt2"))
    fad = FrechetAudioDistance.__new__(FrechetAudioDistance)
    fad.model_name, fad.ckpt_dir, fad.verbose, fad._state_dict = "vggish", str(tmp_path), False, None
    got = fad._resolve_state_dict()
    assert set(got) == set(sd)
    for k in sd:
        assert torch.equal(got[k].detach().cpu(), sd[k]), k
    fad.ckpt_dir = str(tmp_path / "nowhere")
    with pytest.raises(FileNotFoundError):
        fad._resolve_state_dict()
    fad.model_name = "encodec-24k"
    with pytest.raises(NotImplementedError):
        fad._resolve_state_dict()
