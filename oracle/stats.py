"""Oracle (CPU, NumPy/SciPy) restatement of the reference statistics and Frechet distance.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

Reference sites (relative to /root/reference/frechet_audio_distance_exported):
  * calculate_embd_statistics: fad.py:483-496 (mu = np.mean -> float32, sigma = np.cov -> float64, ddof=1)
  * calculate_frechet_distance: fad.py:498-555 (complex Schur sqrtm of sigma1 @ sigma2, eps retry,
    imaginary-diagonal check, trace combination).

scipy >= 1.16 removed the `disp=` keyword the reference passes at fad.py:538, so `_sqrtm` below
is the compat wrapper (same value; the second return of the old API was discarded by the
reference anyway).
"""
from __future__ import annotations

import numpy as np
from scipy import linalg


def embd_statistics(embd):
    """fad.py:492-496."""
    if isinstance(embd, list):
        embd = np.array(embd)
    mu = np.mean(embd, axis=0)
    sigma = np.cov(embd, rowvar=False)
    return mu, sigma


def _sqrtm(a):
    out = linalg.sqrtm(a)
    if isinstance(out, tuple):      # very old scipy with disp=False semantics
        out = out[0]
    return out


def frechet_distance(mu1, sigma1, mu2, sigma2, eps: float = 1e-6):
    """fad.py:525-555."""
    mu1 = np.atleast_1d(mu1)
    mu2 = np.atleast_1d(mu2)
    sigma1 = np.atleast_2d(sigma1)
    sigma2 = np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape, "Training and test mean vectors have different lengths"
    assert sigma1.shape == sigma2.shape, "Training and test covariances have different dimensions"
    diff = mu1 - mu2
    covmean = _sqrtm(sigma1.dot(sigma2).astype(complex))
    if not np.isfinite(covmean).all():
        print("FID calculation produces singular product; adding %s to diagonal of cov estimates" % eps)
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = _sqrtm((sigma1 + offset).dot(sigma2 + offset).astype(complex))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            m = np.max(np.abs(covmean.imag))
            raise ValueError(f"Imaginary component {m}")
        covmean = covmean.real
    tr_covmean = np.trace(covmean)
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * tr_covmean


def frechet_distance_eigh(mu1, sigma1, mu2, sigma2):
    """Symmetric formulation used as a *second* CPU opinion (not the reference algorithm):
    tr sqrtm(S1 S2) = sum_i sqrt(lambda_i(S1^1/2 S2 S1^1/2)), lambda clipped at 0."""
    w, v = np.linalg.eigh(sigma1)
    r = (v * np.sqrt(np.clip(w, 0, None))) @ v.T
    lam = np.linalg.eigvalsh(r @ sigma2 @ r)
    tr = np.sqrt(np.clip(lam, 0, None)).sum()
    diff = np.asarray(mu1, dtype=np.float64) - np.asarray(mu2, dtype=np.float64)
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * tr
