"""Deterministic synthetic clips and embeddings (shard-invariant: keyed by (set, clip index)).

Used by tests, bench.py and make_golden.py to feed IDENTICAL inputs to the oracle and to the B200
path.  This is input generation, not an algorithm restatement; it lives under oracle/ because the
reference's own test-signal helper (tests/test_basic.py:20-24 of the reference: 0.5*sin(2*pi*f*t))
is reproduced here as `sine_clip`.
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0xFAD0


def _rng(set_id: int, clip: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[SEED_BASE + set_id, clip]))


def sine_clip(duration: float, freq: float, sr: int) -> np.ndarray:
    """Reference test signal, tests/test_basic.py:20-24."""
    t = np.linspace(0, duration, int(sr * duration), dtype=np.float32)
    return (np.sin(2 * np.pi * freq * t) * 0.5).astype(np.float32)


def background_clip(clip: int, n: int) -> np.ndarray:
    """white Gaussian noise x U(0.02, 0.3), clipped to [-1, 1], float32."""
    g = _rng(0, clip)
    amp = g.uniform(0.02, 0.3)
    x = g.standard_normal(n, dtype=np.float32) * np.float32(amp)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def eval_clip(clip: int, n: int, sr: int) -> np.ndarray:
    """1/f^alpha noise (alpha ~ U(0.3, 1)), peak-normalised x U(0.05, 0.5), plus a 0.1 sine."""
    g = _rng(1, clip)
    alpha = g.uniform(0.3, 1.0)
    amp = g.uniform(0.05, 0.5)
    f0 = g.uniform(100.0, min(4000.0, 0.45 * sr))
    w = g.standard_normal(n, dtype=np.float32)
    spec = np.fft.rfft(w.astype(np.float64))
    f = np.arange(spec.shape[0], dtype=np.float64)
    f[0] = 1.0
    spec *= f ** (-alpha / 2.0)
    x = np.fft.irfft(spec, n)
    x = x / np.max(np.abs(x)) * amp
    t = np.arange(n, dtype=np.float64) / sr
    x = x + 0.1 * np.sin(2 * np.pi * f0 * t)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def clip_set(set_id: int, first: int, count: int, n: int, sr: int) -> np.ndarray:
    """(count, n) float32; clip index = first + row — same rows whatever the sharding."""
    out = np.empty((count, n), dtype=np.float32)
    for r in range(count):
        out[r] = background_clip(first + r, n) if set_id == 0 else eval_clip(first + r, n, sr)
    return out


def embedding_set(set_id: int, n: int, d: int, seed: int = 0) -> np.ndarray:
    """(n, d) float32 synthetic embeddings for the statistics microbench (BASELINE config 5):
    x = z A + b, z ~ N(0, I), A = randn(d, d)/sqrt(d); set 1 is scaled 1.1 and shifted 0.05."""
    g = np.random.Generator(np.random.Philox(key=[SEED_BASE + 16 + seed, d]))
    a = (g.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)
    b = g.standard_normal(d).astype(np.float32)
    g2 = np.random.Generator(np.random.Philox(key=[SEED_BASE + 32 + seed + set_id, d]))
    z = g2.standard_normal((n, d), dtype=np.float32)
    x = z @ a + b
    if set_id == 1:
        x = x * np.float32(1.1) + np.float32(0.05)
    return np.ascontiguousarray(x.astype(np.float32))
