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
    p3 = str(tmp_path / "f.wav")
    wavfile.write(p3, 44100, a)
    with pytest.raises(NotImplementedError):
        load_audio(p3, 16000, 1)
