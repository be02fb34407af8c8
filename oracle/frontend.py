"""Oracle (CPU, NumPy) restatement of the reference log-mel front ends.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

Reference sites restated here (paths relative to /root/reference/frechet_audio_distance_exported):
  * VGGish: models/vggish.py:17-33 (constants), :102-117 (_frame), :120-122 (_periodic_hann),
    :125-141 (_stft_magnitude), :144-190 (mel matrix), :193-227 (_log_mel_spectrogram),
    :230-279 (waveform_to_examples).
  * PANN / CLAP: models/pann.py:25-59 (PANN_CONFIGS), :68-145 (waveform_to_logmel);
    models/clap.py:41-80 (preprocess_for_clap); fad.py:41-66 (_pad_to_valid_pann_time),
    fad.py:69-91 (_pad_to_clap_time), fad.py:356-359 (CLAP waveform pad).
  * librosa.stft / librosa.filters.mel are NOT in the reference tree (un-vendored, unpinned):
    restated from their documented semantics (centered reflect-padded STFT, Slaney mel scale with
    Slaney area normalisation).  PARITY UNPINNED for these two; cross-checked against
    torch.stft + torchaudio.functional.melscale_fbanks in tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------------------
# VGGish (models/vggish.py:17-33)
# ----------------------------------------------------------------------------------------------
VGGISH_SR = 16000
VGGISH_WIN = 400        # int(round(16000 * 0.025))                      vggish.py:213
VGGISH_HOP = 160        # int(round(16000 * 0.010))                      vggish.py:214
VGGISH_NFFT = 512       # 2 ** ceil(log2(400))                           vggish.py:215
VGGISH_NMEL = 64
VGGISH_FMIN = 125.0
VGGISH_FMAX = 7500.0
VGGISH_LOG_OFFSET = 0.01
VGGISH_PATCH = 96       # frames per example, hop 96 (no overlap)        vggish.py:264-271


def vggish_num_frames(n_samples: int) -> int:
    """vggish.py:113-114 — complete frames only, no padding."""
    if n_samples < VGGISH_WIN:
        return 0
    return 1 + (n_samples - VGGISH_WIN) // VGGISH_HOP


def vggish_num_patches(n_samples: int) -> int:
    """vggish.py:268-271 applied to the (frames, 64) log-mel — trailing frames dropped."""
    f = vggish_num_frames(n_samples)
    if f < VGGISH_PATCH:
        return 0
    return 1 + (f - VGGISH_PATCH) // VGGISH_PATCH


def _htk_mel(hz):
    """vggish.py:144-147 — HTK mel, 1127*ln(1+f/700)."""
    return 1127.0 * np.log(1.0 + np.asarray(hz, dtype=np.float64) / 700.0)


def vggish_mel_matrix() -> np.ndarray:
    """(257, 64) float64 triangular HTK filterbank, DC row zeroed — vggish.py:150-190."""
    nbins = VGGISH_NFFT // 2 + 1
    bin_mel = _htk_mel(np.linspace(0.0, VGGISH_SR / 2.0, nbins))
    edges = np.linspace(_htk_mel(VGGISH_FMIN), _htk_mel(VGGISH_FMAX), VGGISH_NMEL + 2)
    lo, ctr, hi = edges[:-2], edges[1:-1], edges[2:]
    up = (bin_mel[:, None] - lo[None, :]) / (ctr - lo)[None, :]
    down = (hi[None, :] - bin_mel[:, None]) / (hi - ctr)[None, :]
    w = np.maximum(0.0, np.minimum(up, down))
    w[0, :] = 0.0
    return w


def vggish_window() -> np.ndarray:
    """Periodic Hann(400), float64 — vggish.py:120-122."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi / VGGISH_WIN * np.arange(VGGISH_WIN))


_VGGISH_MEL = None


def vggish_logmel(pcm: np.ndarray) -> np.ndarray:
    """(frames, 64) float64 log-mel of a mono 16 kHz clip — vggish.py:193-227.

    Arithmetic stays in float64 exactly like the reference (float32 PCM * float64 window ->
    complex128 rFFT -> magnitude -> float64 matmul -> log).
    """
    global _VGGISH_MEL
    if _VGGISH_MEL is None:
        _VGGISH_MEL = vggish_mel_matrix()
    pcm = np.asarray(pcm)
    if pcm.ndim > 1:                       # vggish.py:245-246
        pcm = pcm.mean(axis=1)
    nf = vggish_num_frames(pcm.shape[0])
    if nf <= 0:
        return np.zeros((0, VGGISH_NMEL), dtype=np.float64)
    idx = np.arange(VGGISH_WIN)[None, :] + VGGISH_HOP * np.arange(nf)[:, None]
    frames = pcm[idx] * vggish_window()[None, :]
    mag = np.abs(np.fft.rfft(frames, VGGISH_NFFT, axis=1))
    return np.log(mag @ _VGGISH_MEL + VGGISH_LOG_OFFSET)


def vggish_examples(pcm: np.ndarray) -> np.ndarray:
    """(patches, 96, 64) float32 — vggish.py:230-279 at the native rate (no resample branch)."""
    lm = vggish_logmel(pcm)
    p = 0 if lm.shape[0] < VGGISH_PATCH else 1 + (lm.shape[0] - VGGISH_PATCH) // VGGISH_PATCH
    if p == 0:
        return np.zeros((0, VGGISH_PATCH, VGGISH_NMEL), dtype=np.float32)
    return lm[: p * VGGISH_PATCH].reshape(p, VGGISH_PATCH, VGGISH_NMEL).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# PANN / CLAP (models/pann.py:25-59)
# ----------------------------------------------------------------------------------------------
PANN_CONFIGS = {
    8000: dict(n_fft=256, hop=80, fmin=50.0, fmax=4000.0),
    16000: dict(n_fft=512, hop=160, fmin=50.0, fmax=8000.0),
    32000: dict(n_fft=1024, hop=320, fmin=50.0, fmax=14000.0),
    48000: dict(n_fft=1024, hop=480, fmin=50.0, fmax=14000.0),   # CLAP, pann.py:51-58
}
PANN_NMEL = 64
CLAP_SR = 48000
CLAP_MAX_SAMPLES = 480000
CLAP_TIME_FRAMES = 1001          # fad.py:38


def _slaney_hz_to_mel(hz):
    hz = np.asarray(hz, dtype=np.float64)
    f_sp = 200.0 / 3.0
    mel = hz / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_part = min_log_mel + np.log(np.maximum(hz, 1e-300) / min_log_hz) / logstep
    return np.where(hz >= min_log_hz, log_part, mel)


def _slaney_mel_to_hz(mel):
    mel = np.asarray(mel, dtype=np.float64)
    f_sp = 200.0 / 3.0
    hz = f_sp * mel
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(mel >= min_log_mel, min_log_hz * np.exp(logstep * (mel - min_log_mel)), hz)


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """(n_mels, n_fft//2+1) float32 — semantics of librosa.filters.mel(htk=False, norm='slaney')
    as called at pann.py:121-127."""
    fft_f = np.arange(n_fft // 2 + 1, dtype=np.float64) * (sr / n_fft)
    mel_pts = np.linspace(_slaney_hz_to_mel(fmin), _slaney_hz_to_mel(fmax), n_mels + 2)
    f = _slaney_mel_to_hz(mel_pts)
    fdiff = np.diff(f)
    ramps = f[:, None] - fft_f[None, :]
    w = np.zeros((n_mels, fft_f.shape[0]), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (f[2:] - f[:-2])
    w *= enorm[:, None]
    return w.astype(np.float32)


def pann_num_frames(n_samples: int, hop: int) -> int:
    """centered STFT: 1 + n // hop."""
    return 1 + n_samples // hop


def pann_padded_frames(t: int) -> int:
    """fad.py:53-59 — smallest 32k-24 >= t."""
    k = (t + 24 + 31) // 32
    v = 32 * k - 24
    if v < t:
        v += 32
    return v


def stft_power_centered(x32: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """(n_fft//2+1, T) float32 power — semantics of librosa.stft(center=True, pad_mode='reflect',
    window='hann', win_length=n_fft) as called at pann.py:107-115, then |.|**2 at pann.py:118.

    FFT of (float64 window * float32 frames) is evaluated in float64 and stored as complex64.
    """
    x32 = np.asarray(x32, dtype=np.float32)
    pad = n_fft // 2
    if x32.shape[0] <= pad:
        raise ValueError("clip shorter than n_fft/2 cannot be reflect-padded")
    xp = np.pad(x32, (pad, pad), mode="reflect")
    t = 1 + (xp.shape[0] - n_fft) // hop
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)      # periodic hann, float64
    idx = np.arange(n_fft)[None, :] + hop * np.arange(t)[:, None]
    spec = np.fft.rfft(xp[idx] * win[None, :], n_fft, axis=1).astype(np.complex64)  # (T, bins)
    mag = np.abs(spec)                                                    # float32
    return (mag ** 2).T


def pann_logmel(pcm: np.ndarray, sr: int) -> np.ndarray:
    """(T, 64) float32 dB log-mel — pann.py:68-145 at the native rate (no resample branch)."""
    cfg = PANN_CONFIGS[sr]
    pcm = np.asarray(pcm)
    if pcm.ndim > 1:                                  # pann.py:96-97
        pcm = pcm.mean(axis=1)
    x = pcm.astype(np.float32)                        # pann.py:104
    power = stft_power_centered(x, cfg["n_fft"], cfg["hop"])
    fb = slaney_mel_filterbank(sr, cfg["n_fft"], PANN_NMEL, cfg["fmin"], cfg["fmax"])
    mel = np.dot(fb, power)                           # float32 sgemm, pann.py:130
    logmel = 10.0 * np.log10(np.maximum(mel, 1e-10))  # pann.py:133-134
    return np.ascontiguousarray(logmel.T.astype(np.float32))


def pann_features(pcm: np.ndarray, sr: int) -> np.ndarray:
    """(T', 64) float32 with zero rows appended so T' = 32k-24 — fad.py:41-66,377."""
    lm = pann_logmel(pcm, sr)
    tp = pann_padded_frames(lm.shape[0])
    out = np.zeros((tp, PANN_NMEL), dtype=np.float32)
    out[: lm.shape[0]] = lm
    return out


def clap_quantize(pcm: np.ndarray) -> np.ndarray:
    """clap.py:70-72 — int16 truncation toward zero."""
    x = np.asarray(pcm).astype(np.float32)
    return (x * np.float32(32767.0)).astype(np.int16).astype(np.float32) / np.float32(32767.0)


def clap_features(pcm: np.ndarray) -> np.ndarray:
    """(1001, 64) float32 — fad.py:351-362 + clap.py:41-80."""
    pcm = np.asarray(pcm)
    if pcm.ndim > 1:
        pcm = pcm.mean(axis=1)
    if pcm.shape[0] < CLAP_MAX_SAMPLES:               # fad.py:356-359
        pcm = np.pad(pcm, (0, CLAP_MAX_SAMPLES - pcm.shape[0]))
    lm = pann_logmel(clap_quantize(pcm), CLAP_SR)
    out = np.zeros((CLAP_TIME_FRAMES, PANN_NMEL), dtype=np.float32)   # fad.py:81-89
    t = min(CLAP_TIME_FRAMES, lm.shape[0])
    out[:t] = lm[:t]
    return out
