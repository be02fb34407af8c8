"""Thin torch-facing wrapper over the C ABI (include/fadb.h).  PyTorch is only plumbing here: device
memory, streams and torch.distributed; all arithmetic is in libfadb200.so.

One `Engine` = one fadb_handle on one GPU with one model's weights committed.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import MODEL_IDS, PREC_IDS, PREC_PACKING, check

SAMPLE_RATES = {"vggish": 16000, "pann-8k": 8000, "pann-16k": 16000, "pann-32k": 32000, "clap": 48000}
EMBED_DIMS = {"vggish": 128, "pann-8k": 2048, "pann-16k": 2048, "pann-32k": 2048, "clap": 512}


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("frechet_audio_distance_exported_b200 needs a B200 GPU: there is no CPU fallback")


def _stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    def __init__(self, model_name: str, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 precision: str = "fp16x2", device: Optional[int] = None, max_batch: Optional[int] = None):
        if model_name not in MODEL_IDS:
            raise ValueError(f"Unknown model: {model_name}. Valid options: {list(MODEL_IDS)}")
        _require_cuda()
        self.lib = _lib.load()
        self.model_name = model_name
        self.model_id = MODEL_IDS[model_name]
        self.sample_rate = SAMPLE_RATES[model_name]
        self.dim = EMBED_DIMS[model_name]
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.handle = _lib.Handle(self.device_index)
        self.h = self.handle.ptr
        self.weights_loaded = False
        self._sd = None
        self._packed = None
        self.set_precision(precision)
        self.max_batch = 16384                 # items per network launch (csrc/common.cuh default)
        if max_batch is not None:
            check(self.lib.fadb_set_max_batch(self.h, int(max_batch)))
            self.max_batch = int(max_batch)
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # ------------------------------------------------------------------ configuration
    def set_precision(self, precision: str) -> None:
        if precision not in PREC_IDS:
            raise ValueError(f"precision must be one of {list(PREC_IDS)}")
        check(self.lib.fadb_set_precision(self.h, PREC_IDS[precision]))
        self.precision = precision
        # the packed weights are in the 16-bit format (bf16 / fp16, with or without the lo plane) of the precision
        # they were committed under: re-pack when the new mode needs something else
        fmt, lo = PREC_PACKING[precision]
        if self.weights_loaded and (self._packed[0] != fmt or (lo and not self._packed[1])):
            self.load_state_dict(self._sd)

    def set_tensor_syrk(self, on: bool) -> None:
        """Second moments of large sets (>= 8192 rows, d >= 512) on the tensor cores: ~8x faster, covariance to ~1e-6
        instead of fp64-exact.  Off by default (include/fadb.h: fadb_set_tensor_syrk)."""
        check(self.lib.fadb_set_tensor_syrk(self.h, int(bool(on))))

    def set_clap_quantize(self, on: bool) -> None:
        """CLAP front end: apply clap.py:70-72's int16 truncation to the samples (default) or take them as they are
        (already quantised at the source rate by the caller, include/fadb.h: fadb_set_clap_quantize)."""
        check(self.lib.fadb_set_clap_quantize(self.h, int(bool(on))))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Hand the reference modules' state_dict (VGGishCore / PANNCore key names) to the packer."""
        check(self.lib.fadb_weights_begin(self.h, self.model_id))
        for name, t in sd.items():
            if not torch.is_tensor(t) or not t.is_floating_point():
                continue                                  # num_batches_tracked etc.
            a = t.detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * max(a.dim(), 1))(*a.shape)
            check(self.lib.fadb_weights_tensor(self.h, name.encode(), C.c_void_p(a.data_ptr()), shape, a.dim()))
        check(self.lib.fadb_weights_commit(self.h))
        self.weights_loaded = True
        self._sd = sd
        self._packed = PREC_PACKING[self.precision]

    # ------------------------------------------------------------------ hot path pieces (device tensors)
    def frontend_rows(self, n_samples: int) -> int:
        return int(self.lib.fadb_frontend_rows(self.model_id, int(n_samples)))

    def frontend(self, pcm: torch.Tensor) -> torch.Tensor:
        """pcm [n_clips, n_samples] fp32 cuda -> VGGish [n_clips*P, 96, 64] / CNN14 [n_clips, T', 64] fp32."""
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.dim() == 2 and pcm.stride(1) == 1
        n_clips, n = pcm.shape
        rows = self.frontend_rows(n)
        if self.model_id == 0:
            out = torch.empty((n_clips * max(rows, 0), 96, 64), dtype=torch.float32, device=pcm.device)
        else:
            out = torch.empty((n_clips, rows, 64), dtype=torch.float32, device=pcm.device)
        if out.numel():
            check(self.lib.fadb_frontend(self.h, self.model_id, _p(pcm), n_clips, n, pcm.stride(0), _p(out),
                                         C.c_void_p(_stream_ptr())))
        return out

    def embed_features(self, feats: torch.Tensor) -> torch.Tensor:
        """feats [items, T, 64] fp32 cuda -> [items, d] fp32 (the reference's `self.model(x)`)."""
        assert feats.is_cuda and feats.dtype == torch.float32 and feats.dim() == 3 and feats.shape[2] == 64
        feats = feats.contiguous()
        out = torch.empty((feats.shape[0], self.dim), dtype=torch.float32, device=feats.device)
        if feats.shape[0]:
            check(self.lib.fadb_embed(self.h, _p(feats), feats.shape[0], feats.shape[1], _p(out),
                                      C.c_void_p(_stream_ptr())))
        return out

    def embed_pcm(self, pcm: torch.Tensor) -> torch.Tensor:
        """pcm [n_clips, n_samples] cuda, fp32 in [-1, 1] or raw int16 PCM (scaled by 1/32768 in the front end)
        -> embeddings [rows, d] fp32 (front end + network)."""
        assert pcm.is_cuda and pcm.dtype in (torch.float32, torch.int16) and pcm.dim() == 2 and pcm.stride(1) == 1
        n_clips, n = pcm.shape
        rows = self.frontend_rows(n) if self.model_id == 0 else 1
        out = torch.empty((n_clips * max(rows, 0), self.dim), dtype=torch.float32, device=pcm.device)
        if out.numel():
            fn = self.lib.fadb_embed_pcm16 if pcm.dtype == torch.int16 else self.lib.fadb_embed_pcm
            check(fn(self.h, _p(pcm), n_clips, n, pcm.stride(0), _p(out), C.c_void_p(_stream_ptr())))
        return out

    def resample(self, pcm: torch.Tensor, sr_orig: int, sr_new: int) -> torch.Tensor:
        """pcm [n_clips, n] fp32 cuda at sr_orig -> [n_clips, int(n * sr_new / sr_orig)] fp32 at sr_new, with the
        semantics (and the exact fp64 arithmetic) of resample.resample / resampy kaiser_best."""
        from .resample import filter_table
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.dim() == 2 and pcm.stride(1) == 1
        key = (int(sr_orig), int(sr_new))
        if not hasattr(self, "_resample_tables"):
            self._resample_tables = {}
        if key not in self._resample_tables:
            win, num_table, ratio = filter_table(*key)
            self._resample_tables[key] = (torch.from_numpy(win).to(self.device), num_table, ratio)
        win_dev, num_table, ratio = self._resample_tables[key]
        n_clips, n = pcm.shape
        n_out = int(n * ratio)
        if n_out < 1:
            raise ValueError(f"Input signal length={n} is too small to resample from {sr_orig}->{sr_new}")
        out = torch.empty((n_clips, n_out), dtype=torch.float32, device=pcm.device)
        if n_clips:
            check(self.lib.fadb_resample(self.h, _p(pcm), n_clips, n, pcm.stride(0), C.c_double(ratio), _p(win_dev),
                                         win_dev.numel(), num_table, _p(out), n_out, out.stride(0),
                                         C.c_void_p(_stream_ptr())))
        return out

    # ------------------------------------------------------------------ statistics
    def new_acc(self, d: Optional[int] = None) -> torch.Tensor:
        d = self.dim if d is None else d
        return torch.zeros(1 + d + d * d, dtype=torch.float64, device=self.device)

    def stats_accumulate(self, emb: torch.Tensor, acc: torch.Tensor, shift: Optional[torch.Tensor] = None) -> None:
        assert emb.is_cuda and emb.dim() == 2 and emb.stride(1) == 1
        d = emb.shape[1]
        assert acc.numel() == 1 + d + d * d and acc.dtype == torch.float64
        if emb.shape[0] == 0:
            return
        fn = self.lib.fadb_stats_accumulate if emb.dtype == torch.float32 else self.lib.fadb_stats_accumulate_f64
        assert emb.dtype in (torch.float32, torch.float64)
        check(fn(self.h, _p(emb), emb.shape[0], d, emb.stride(0), _p(shift), _p(acc), C.c_void_p(_stream_ptr())))

    def stats_finalize(self, acc: torch.Tensor, d: int, shift: Optional[torch.Tensor] = None):
        mu = torch.empty(d, dtype=torch.float64, device=acc.device)
        sigma = torch.empty((d, d), dtype=torch.float64, device=acc.device)
        check(self.lib.fadb_stats_finalize(self.h, _p(acc), d, _p(shift), _p(mu), _p(sigma), C.c_void_p(_stream_ptr())))
        return mu, sigma

    def allreduce_acc(self, acc: torch.Tensor, group=None) -> None:
        """The only collective of the path: one NCCL all-reduce (sum, fp64) of {n, sum x, sum x x^T}."""
        from .dist import allreduce_acc
        allreduce_acc(acc, group)

    # ------------------------------------------------------------------ Frechet
    def frechet(self, mu1, sigma1, mu2, sigma2) -> torch.Tensor:
        """device fp64 tensors -> device fp64[4] = {fad, tr sqrt(S1 S2), tr S1 + tr S2, |mu1-mu2|^2}."""
        d = mu1.numel()
        for t in (mu1, sigma1, mu2, sigma2):
            assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        out = torch.empty(4, dtype=torch.float64, device=mu1.device)
        check(self.lib.fadb_frechet(self.h, _p(mu1), _p(sigma1), _p(mu2), _p(sigma2), d, _p(out),
                                    C.c_void_p(_stream_ptr())))
        return out

    # ------------------------------------------------------------------ whole path, host buffers
    def fad_from_pcm_host(self, pcm_bg: torch.Tensor, pcm_ev: torch.Tensor, return_embeddings: bool = False):
        """pcm_* : [n_clips, n_samples] fp32 (or both raw int16 PCM) HOST tensors (pinned recommended).  One C call:
        chunked H2D overlapped with compute, statistics, Frechet, scalar D2H."""
        for t in (pcm_bg, pcm_ev):
            assert (not t.is_cuda) and t.dtype in (torch.float32, torch.int16) and t.dim() == 2 and t.is_contiguous()
        assert pcm_bg.shape[1] == pcm_ev.shape[1] and pcm_bg.dtype == pcm_ev.dtype
        n = pcm_bg.shape[1]
        rows = self.frontend_rows(n) if self.model_id == 0 else 1
        eb = ee = None
        if return_embeddings:
            eb = torch.empty((pcm_bg.shape[0] * rows, self.dim), dtype=torch.float32).pin_memory()
            ee = torch.empty((pcm_ev.shape[0] * rows, self.dim), dtype=torch.float32).pin_memory()
        out = C.c_double(0.0)
        fn = self.lib.fadb_fad_from_pcm16_host if pcm_bg.dtype == torch.int16 else self.lib.fadb_fad_from_pcm_host
        check(fn(self.h, _p(pcm_bg), pcm_bg.shape[0], _p(pcm_ev), pcm_ev.shape[0], n, _p(eb), _p(ee), C.byref(out)))
        if return_embeddings:
            return out.value, eb.numpy(), ee.numpy()
        return out.value

    def profile_enable(self, on: bool = True) -> None:
        check(self.lib.fadb_profile_enable(self.h, int(on)))

    def profile_read(self):
        """-> (tensor-core layer ms, algorithmic FLOPs, launches) since profile_enable(True)."""
        out = (C.c_double * 4)()
        check(self.lib.fadb_profile_read(self.h, out))
        self.front_ms = float(out[3])          # summed front-end (+ fused conv1) launch durations of the same window
        return float(out[0]), float(out[1]), int(out[2])

    def launch_count(self) -> int:
        return self.handle.launch_count()

    def device_status(self) -> int:
        return int(self.lib.fadb_device_status(self.h))

    # ------------------------------------------------------------------ test hook
    def debug_conv_layer(self, x_nhwc: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], ksize: int,
                         relu: bool, pool: int) -> torch.Tensor:
        B, H, W, Cin = x_nhwc.shape
        Cout = w.shape[0]
        Ho, Wo = (H // 2, W // 2) if pool else (H, W)
        out = torch.zeros((B, Ho, Wo, Cout), dtype=torch.float32, device=x_nhwc.device)
        check(self.lib.fadb_debug_conv_layer(self.h, _p(x_nhwc.contiguous()), B, H, W, Cin, _p(w.contiguous()),
                                             _p(bias), Cout, ksize, int(relu), int(pool), _p(out),
                                             C.c_void_p(_stream_ptr())))
        return out

