"""CPU oracle for the FAD hot path — TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy / torch-CPU, the arithmetic of the reference
`frechet_audio_distance_exported` hot path (PCM -> log-mel -> VGGish / CNN14
embedding -> mean/cov -> Frechet distance).  Every function cites the reference
file:line it follows.

Rules (enforced by tests/test_abi.py::test_product_never_imports_the_oracle):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
    `--impl reference` legs may import anything from here;
  * the product package `frechet_audio_distance_exported_b200` never imports it and has no
    CPU fallback — it fails loudly when the CUDA library is missing.

Parity pinning:
  * VGGish front end, VGGishCore, PANNCore, statistics, Frechet: PINNED — checked
    against the reference's own code imported through `oracle/ref_shim.py` (script
    `oracle/make_golden.py`; fixtures in `tests/golden/`), and against the
    reference's known-answer tests (`tests/test_basic.py:143-170` of the reference).
  * PANN / CLAP front end: the arithmetic lives in un-vendored `librosa`
    (no version pin in the reference, not installed here) — PARITY UNPINNED for
    `librosa.stft` / `librosa.filters.mel`; restated from their published
    semantics and cross-checked against torch.stft + torchaudio's Slaney
    filterbank (independent second implementation).
  * CLAP CNN14 head: no reference code exists (README only) — PARITY UNPINNED.
"""
