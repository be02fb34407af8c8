"""Import shim that makes the UNMODIFIED reference package importable in the builder container.

TEST INFRASTRUCTURE ONLY.  Only works where /root/reference exists (the builder container); the
GPU box never uses it — the outputs it produced are committed under tests/golden/ by
oracle/make_golden.py.

What is stubbed and why:
  * `resampy`  — not installed; native-rate inputs never reach it (vggish.py:249, pann.py:100).
                 The stub raises if it is ever called.
  * `soundfile` — not installed; only the file-based reference tests need it.  PCM16 WAV via `wave`.
  * `librosa`  — not installed; `stft` and `filters.mel` are provided by oracle/frontend.py's
                 restatement (=> the PANN/CLAP front end stays PARITY UNPINNED).
  * `scipy.linalg.sqrtm(disp=...)` — keyword removed in scipy 1.16+, the reference passes it
                 (fad.py:538); wrapped to the old semantics.
"""
from __future__ import annotations

import os
import sys
import types
import wave

import numpy as np

REFERENCE_ROOT = os.environ.get("FADB_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "frechet_audio_distance_exported"))


def _install_stubs():
    from . import frontend

    if "resampy" not in sys.modules:
        m = types.ModuleType("resampy")

        def resample(*a, **k):
            raise RuntimeError("resampy is not installed (oracle shim): non-native sample rate reached")

        m.resample = resample
        sys.modules["resampy"] = m

    if "soundfile" not in sys.modules:
        m = types.ModuleType("soundfile")

        def write(fname, data, sr, subtype="PCM_16"):
            data = np.asarray(data)
            if data.ndim == 1:
                data = data[:, None]
            pcm = np.clip(np.round(data * 32768.0), -32768, 32767).astype("<i2")
            with wave.open(fname, "wb") as w:
                w.setnchannels(pcm.shape[1])
                w.setsampwidth(2)
                w.setframerate(int(sr))
                w.writeframes(pcm.tobytes())

        def read(fname, dtype="float32"):
            with wave.open(fname, "rb") as w:
                ch, sr, n = w.getnchannels(), w.getframerate(), w.getnframes()
                raw = np.frombuffer(w.readframes(n), dtype="<i2").reshape(-1, ch)
            if dtype == "int16":
                out = raw.astype(np.int16)
            elif dtype == "int32":
                out = raw.astype(np.int32) << 16
            else:
                out = (raw.astype(np.float64) / 32768.0).astype(dtype)
            if ch == 1:
                out = out[:, 0]
            return out, sr

        m.read, m.write = read, write
        sys.modules["soundfile"] = m

    if "librosa" not in sys.modules:
        m = types.ModuleType("librosa")
        filt = types.ModuleType("librosa.filters")

        def stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
                 pad_mode="reflect"):
            assert window == "hann" and center and pad_mode == "reflect"
            assert win_length in (None, n_fft)
            y = np.asarray(y, dtype=np.float32)
            pad = n_fft // 2
            yp = np.pad(y, (pad, pad), mode="reflect")
            t = 1 + (yp.shape[0] - n_fft) // hop_length
            win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)
            idx = np.arange(n_fft)[None, :] + hop_length * np.arange(t)[:, None]
            return np.fft.rfft(yp[idx] * win[None, :], n_fft, axis=1).astype(np.complex64).T

        def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
            return frontend.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax if fmax else sr / 2)

        m.stft = stft
        filt.mel = mel
        m.filters = filt
        sys.modules["librosa"] = m
        sys.modules["librosa.filters"] = filt

    import scipy.linalg as sl

    if not getattr(sl.sqrtm, "_fadb_wrapped", False):
        real = sl.sqrtm

        def sqrtm(a, disp=True, blocksize=None):
            x = real(a)
            if disp:
                return x
            a = np.asarray(a)
            err = np.linalg.norm(x @ x - a, "fro") / max(np.linalg.norm(a, "fro"), 1e-300)
            return x, err

        sqrtm._fadb_wrapped = True
        sl.sqrtm = sqrtm


def load_reference():
    """Returns the imported `frechet_audio_distance_exported` package from /root/reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import frechet_audio_distance_exported as ref  # noqa

    return ref


def make_reference_fad(ref, model_name: str, module):
    """Reference FrechetAudioDistance instance without download (same trick as the reference's
    tests/test_basic.py:136-141), with `module` as `.model` on CPU."""
    import torch

    fad = ref.fad.FrechetAudioDistance.__new__(ref.fad.FrechetAudioDistance)
    fad.model_name = model_name
    fad.sample_rate = ref.fad.VALID_MODELS[model_name]["sample_rate"]
    fad.channels = 1
    fad.verbose = False
    fad.audio_load_worker = 1
    fad.device = torch.device("cpu")
    fad.model = module
    return fad
