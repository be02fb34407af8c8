"""ctypes binding of libfadb200.so (C ABI in include/fadb.h).

There is NO CPU fallback: if the library cannot be built/loaded, or no B200 is present, the calls
raise.  Loading the library itself does not need a GPU (symbol checks work on the build box).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FADB_LIB_PATH") or os.path.join(_HERE, "libfadb200.so")   # override: A/B profiling only

MODEL_IDS = {"vggish": 0, "pann-8k": 1, "pann-16k": 2, "pann-32k": 3, "clap": 4}
# arithmetic of the tensor-core layers (include/fadb.h FADB_PREC_*); "fp16x2" is the library default
PREC_IDS = {"bf16": 0, "bf16x3": 1, "fp16": 2, "fp16x2": 3}
# what fadb_weights_commit packs for a precision: (16-bit format, lo plane present)
PREC_PACKING = {"bf16": ("bf16", False), "bf16x3": ("bf16", True), "fp16": ("fp16", False), "fp16x2": ("fp16", True)}

# every symbol include/fadb.h declares: (name, restype, argtypes)
_vp, _i64, _i32, _fp, _dp = C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p
SYMBOLS = [
    ("fadb_abi_version", C.c_int, []),
    ("fadb_last_error", C.c_char_p, []),
    ("fadb_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    ("fadb_destroy", None, [_vp]),
    ("fadb_set_precision", C.c_int, [_vp, C.c_int]),
    ("fadb_set_max_batch", C.c_int, [_vp, C.c_int]),
    ("fadb_set_tensor_syrk", C.c_int, [_vp, C.c_int]),
    ("fadb_set_clap_quantize", C.c_int, [_vp, C.c_int]),
    ("fadb_weights_begin", C.c_int, [_vp, C.c_int]),
    ("fadb_weights_tensor", C.c_int, [_vp, C.c_char_p, _fp, C.POINTER(C.c_int64), C.c_int]),
    ("fadb_weights_commit", C.c_int, [_vp]),
    ("fadb_frontend_rows", C.c_int64, [C.c_int, _i64]),
    ("fadb_frontend", C.c_int, [_vp, C.c_int, _fp, _i64, _i64, _i64, _fp, _vp]),
    ("fadb_embed_dim", C.c_int, [C.c_int]),
    ("fadb_embed", C.c_int, [_vp, _fp, _i64, _i64, _fp, _vp]),
    ("fadb_resample", C.c_int, [_vp, _fp, _i64, _i64, _i64, C.c_double, _dp, C.c_int, C.c_int, _fp, _i64, _i64, _vp]),
    ("fadb_embed_pcm", C.c_int, [_vp, _fp, _i64, _i64, _i64, _fp, _vp]),
    ("fadb_embed_pcm16", C.c_int, [_vp, _vp, _i64, _i64, _i64, _fp, _vp]),
    ("fadb_stats_accumulate", C.c_int, [_vp, _fp, _i64, C.c_int, _i64, _dp, _dp, _vp]),
    ("fadb_stats_accumulate_f64", C.c_int, [_vp, _dp, _i64, C.c_int, _i64, _dp, _dp, _vp]),
    ("fadb_stats_finalize", C.c_int, [_vp, _dp, C.c_int, _dp, _dp, _dp, _vp]),
    ("fadb_frechet", C.c_int, [_vp, _dp, _dp, _dp, _dp, C.c_int, _dp, _vp]),
    ("fadb_fad_from_pcm_host", C.c_int, [_vp, _fp, _i64, _fp, _i64, _i64, _fp, _fp, C.POINTER(C.c_double)]),
    ("fadb_fad_from_pcm16_host", C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _fp, _fp, C.POINTER(C.c_double)]),
    ("fadb_profile_enable", C.c_int, [_vp, C.c_int]),
    ("fadb_profile_read", C.c_int, [_vp, C.POINTER(C.c_double)]),
    ("fadb_launch_count", C.c_int64, [_vp]),
    ("fadb_device_status", C.c_int, [_vp]),
    ("fadb_debug_conv_layer", C.c_int,
     [_vp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _vp]),
]

_lib = None


class FadbError(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libfadb200.so (building it with nvcc first if it is not there) and bind every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise FadbError(f"{LIB_PATH} is missing; run `python -m frechet_audio_distance_exported_b200.build` "
                            "(no CPU fallback exists)")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.fadb_abi_version() != 1:
        raise FadbError("libfadb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().fadb_last_error()
        raise FadbError(f"libfadb200 error {rc}: {msg.decode() if msg else '?'}")


class Handle:
    """RAII wrapper of fadb_handle*."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self._h = C.c_void_p()
        check(self.lib.fadb_create(C.byref(self._h), int(device)))
        self.device = int(device)

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._h = None
            try:
                self.lib.fadb_destroy(h)
            except Exception:          # interpreter shutdown: ctypes globals may already be gone
                pass

    __del__ = close

    @property
    def ptr(self):
        return self._h

    def launch_count(self) -> int:
        return int(self.lib.fadb_launch_count(self._h))
