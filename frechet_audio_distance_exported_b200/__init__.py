"""frechet_audio_distance_exported_b200 — B200-native (sm_100a) hot path of
gibiansky/frechet-audio-distance-exported behind the reference's own Python surface.

    from frechet_audio_distance_exported_b200 import FrechetAudioDistance

The arithmetic lives in libfadb200.so (hand-written CUDA, C ABI in include/fadb.h).  There is no CPU
fallback: without the library or without a B200 the calls raise.
"""
from .fad import (  # noqa: F401
    CLAP_TIME_FRAMES,
    ENCODEC_SAMPLE_RATES,
    EXPORTED_MODEL_URLS,
    PANN_SAMPLE_RATES,
    VALID_MODELS,
    FrechetAudioDistance,
    _pad_to_clap_time,
    _pad_to_valid_pann_time,
    load_audio,
)
from .engine import Engine  # noqa: F401

__version__ = "0.1.0"
__all__ = ["FrechetAudioDistance", "Engine", "load_audio", "VALID_MODELS"]
