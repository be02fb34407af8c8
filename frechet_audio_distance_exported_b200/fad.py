"""Drop-in façade: same class, method names, argument meaning and error behaviour as the reference
`frechet_audio_distance_exported.fad.FrechetAudioDistance` (fad.py:164-662), with the hot path
(PCM -> log-mel -> VGGish / CNN14 embedding -> mean/cov -> Frechet) running in libfadb200.so on a
B200.  There is no CPU fallback.

Differences a user of the reference sees:
  * weights: the reference downloads `*.pt2` artefacts (fad.py:249-300); offline we accept a
    state_dict of the reference's own modules (`state_dict=` or a file in `ckpt_dir`);
  * "clap" is the CNN14 audio branch + projection head (BASELINE.json / README of the reference),
    not the HTSAT artefact; "encodec-*" is out of scope and raises NotImplementedError;
  * resampling follows resampy's published kaiser_best algorithm (resample.py, host side like the
    reference); resampy itself is un-vendored and absent here, so this piece is parity-unpinned.
"""
from __future__ import annotations

import os
from multiprocessing.dummy import Pool as ThreadPool
from typing import Dict, List, Optional

import numpy as np
import torch

from .engine import Engine
from .resample import resample
from .stream import HostRing

CLAP_TIME_FRAMES = 1001                       # fad.py:38
_CLAP_MAX_SAMPLES = 480000                    # fad.py:355 (CLAP_MAX_SECONDS * CLAP_SAMPLE_RATE)

# fad.py:109-117
VALID_MODELS = {
    "vggish": {"sample_rate": 16000, "embedding_dim": 128},
    "pann-8k": {"sample_rate": 8000, "embedding_dim": 2048},
    "pann-16k": {"sample_rate": 16000, "embedding_dim": 2048},
    "pann-32k": {"sample_rate": 32000, "embedding_dim": 2048},
    "encodec-24k": {"sample_rate": 24000, "embedding_dim": 128, "channels": 1},
    "encodec-48k": {"sample_rate": 48000, "embedding_dim": 128, "channels": 2},
    "clap": {"sample_rate": 48000, "embedding_dim": 512},
}
PANN_SAMPLE_RATES = {"pann-8k": 8000, "pann-16k": 16000, "pann-32k": 32000}        # fad.py:120-124
ENCODEC_SAMPLE_RATES = {"encodec-24k": 24000, "encodec-48k": 48000}               # fad.py:127-130
# fad.py:95-106 — kept for API compatibility; nothing is downloaded (no network on the GPU boxes)
EXPORTED_MODEL_URLS = {
    "vggish": "https://github.com/gibiansky/frechet-audio-distance-exported/releases/download/v0.1/vggish_exported.pt2",
    "pann-8k": "https://github.com/gibiansky/frechet-audio-distance-exported/releases/download/v0.2/pann_cnn14_8k_exported.pt2",
    "pann-16k": "https://github.com/gibiansky/frechet-audio-distance-exported/releases/download/v0.2/pann_cnn14_16k_exported.pt2",
    "pann-32k": "https://github.com/gibiansky/frechet-audio-distance-exported/releases/download/v0.2/pann_cnn14_32k_exported.pt2",
    "clap": "https://github.com/gibiansky/frechet-audio-distance-exported/releases/download/v0.3/clap_exported.pt2",
}


def _pad_to_valid_pann_time(x: torch.Tensor) -> torch.Tensor:
    """fad.py:41-66 — kept for callers that feed `model(x)` themselves; the fused front end already
    emits the padded layout."""
    time = x.shape[2]
    k = (time + 24 + 31) // 32
    valid_time = 32 * k - 24
    if valid_time < time:
        valid_time += 32
    if valid_time > time:
        x = torch.nn.functional.pad(x, (0, 0, 0, valid_time - time))
    return x


def _pad_to_clap_time(x: torch.Tensor) -> torch.Tensor:
    """fad.py:69-91."""
    time = x.shape[2]
    if time < CLAP_TIME_FRAMES:
        x = torch.nn.functional.pad(x, (0, 0, 0, CLAP_TIME_FRAMES - time))
    elif time > CLAP_TIME_FRAMES:
        x = x[:, :, :CLAP_TIME_FRAMES, :]
    return x


class RawPCM16(np.ndarray):
    """int16 samples straight from a 16-bit WAV file, still to be divided by 32768 (fad.py:148-149).  Only load_audio
    creates it; a plain int16 ndarray handed to get_embeddings is NOT rescaled, like in the reference."""


def load_audio(fname: str, sample_rate: int, channels: int, dtype: str = "float32", raw_pcm16: bool = False) -> np.ndarray:
    """fad.py:133-161 for RIFF/WAV files (PCM 8/16/24/32-bit and IEEE float), without libsndfile.
    raw_pcm16=True (not in the reference): a mono 16-bit file already at `sample_rate` is returned as the raw
    int16 samples; get_embeddings ships them to the GPU as they are and the front end applies the /32768 of
    fad.py:148-149 there (same values, half the bytes)."""
    from scipy.io import wavfile

    sr, raw = wavfile.read(fname)
    if raw_pcm16 and raw.dtype == np.int16 and raw.ndim == 1 and sr == sample_rate:
        return raw.view(RawPCM16)
    if raw.dtype == np.uint8:
        f = (raw.astype(np.float64) - 128.0) / 128.0
    elif raw.dtype == np.int16:
        f = raw.astype(np.float64) / 32768.0
    elif raw.dtype == np.int32:                      # 24-bit is left-justified in int32 by scipy
        f = raw.astype(np.float64) / 2147483648.0
    else:
        f = raw.astype(np.float64)
    if dtype == "int16":                              # fad.py:148-149 (sf.read int16, then / 32768.0)
        wav_data = np.clip(np.round(f * 32768.0), -32768, 32767).astype(np.int16) / 32768.0
    elif dtype == "int32":                            # fad.py:150-151
        wav_data = np.clip(np.round(f * 2147483648.0), -2147483648, 2147483647).astype(np.int32) / float(2 ** 31)
    else:
        wav_data = f.astype(dtype)
    if len(wav_data.shape) > channels:                # fad.py:154-155
        wav_data = np.mean(wav_data, axis=1)
    if sr != sample_rate:                             # fad.py:158-159
        wav_data = resample(wav_data, sr, sample_rate)
    return wav_data


class _B200Model:
    """Callable with the contract of the reference's `self.model` (fad.py:297-299):
    [B,1,96,64] -> [B,128]; [B,1,T,64] -> [B,2048]; [B,1,1001,64] -> [B,512]; fp32 torch tensors."""

    def __init__(self, engine: Engine):
        self.engine = engine

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] != 1 or x.shape[3] != 64:
            raise ValueError(f"expected [B,1,T,64], got {tuple(x.shape)}")
        x = x.to(self.engine.device, torch.float32)
        return self.engine.embed_features(x[:, 0].contiguous())

    def to(self, *a, **k):
        return self

    def eval(self):
        return self


_GATHER_POOL = None


def _gather_clips(view: np.ndarray, clips: List[np.ndarray], c0: int, nc: int) -> None:
    """Copy clips[c0 : c0 + nc] into the rows of a pinned staging buffer.  One thread copies ~6-10 GB/s = 10-15 k
    ten-second clips/s, a quarter of what one GPU embeds; numpy releases the GIL in the copy, so four threads share it."""
    global _GATHER_POOL
    workers = min(4, os.cpu_count() or 1)
    if nc < 16 or workers < 2:
        for j in range(nc):
            view[j] = clips[c0 + j]
        return
    if _GATHER_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _GATHER_POOL = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="fadb-gather")

    def part(w: int) -> None:
        for j in range(w, nc, workers):
            view[j] = clips[c0 + j]

    list(_GATHER_POOL.map(part, range(workers)))


class FrechetAudioDistance:
    """API-compatible FAD calculator (reference fad.py:164-662) on libfadb200.so."""

    def __init__(
        self,
        ckpt_dir: Optional[str] = None,
        model_name: str = "vggish",
        sample_rate: Optional[int] = None,
        channels: int = 1,
        verbose: bool = False,
        audio_load_worker: int = 8,
        *,
        state_dict: Optional[Dict[str, torch.Tensor]] = None,
        precision: str = "fp16x2",
        process_group=None,
    ):
        if model_name not in VALID_MODELS:                                     # fad.py:205-208
            raise ValueError(f"Unknown model: {model_name}. Valid options: {list(VALID_MODELS.keys())}")
        expected_sr = VALID_MODELS[model_name]["sample_rate"]
        if sample_rate is None:                                                # fad.py:214-219
            sample_rate = expected_sr
        elif sample_rate != expected_sr:
            raise ValueError(f"Model '{model_name}' requires sample_rate={expected_sr}, got {sample_rate}")
        self.model_name = model_name
        self.sample_rate = sample_rate
        self.channels = channels
        self.verbose = verbose
        self.audio_load_worker = audio_load_worker
        self.precision = precision
        self.process_group = process_group
        self.device = torch.device("cuda")                                     # fad.py:228-233: CUDA only, no mps/cpu branch
        if self.verbose:
            print(f"[Exported FAD] Using device: {self.device}")
        if ckpt_dir is not None:                                               # fad.py:239-244
            os.makedirs(ckpt_dir, exist_ok=True)
            self.ckpt_dir = ckpt_dir
        else:
            self.ckpt_dir = os.path.join(torch.hub.get_dir(), "exported_fad")
            os.makedirs(self.ckpt_dir, exist_ok=True)
        self._state_dict = state_dict
        self._load_model()

    # ------------------------------------------------------------------ fad.py:249-300
    def _model_filename(self) -> str:
        if self.model_name == "vggish":
            return "vggish_exported.pt2"
        if self.model_name in PANN_SAMPLE_RATES:
            return f"pann_cnn14_{self.model_name.split('-')[1]}_exported.pt2"
        if self.model_name == "clap":
            return "clap_exported.pt2"
        return f"{self.model_name}_exported.pt2"

    def _resolve_state_dict(self) -> Dict[str, torch.Tensor]:
        """Host-side half of fad.py:249-300: find the weights.  `state_dict=` wins; else the exported artefact
        `<ckpt_dir>/<name>_exported.pt2` (torch.export, read for its state_dict only — the graph is not run:
        the network executes in libfadb200.so) or `<...>_state_dict.pt`.  Nothing is downloaded."""
        if self.model_name in ENCODEC_SAMPLE_RATES:
            raise NotImplementedError("Encodec is out of scope of the B200 hot path (no model source in the reference)")
        if self._state_dict is not None:
            return self._state_dict
        path = os.path.join(self.ckpt_dir, self._model_filename())
        alt = os.path.splitext(path)[0] + "_state_dict.pt"
        if os.path.exists(path):
            if self.verbose:
                print(f"[Exported FAD] Loading model from {path}...")
            return dict(torch.export.load(path).state_dict)                    # fad.py:297
        if os.path.exists(alt):
            return torch.load(alt, map_location="cpu")
        raise FileNotFoundError(
            f"Exported model not found at {path} (or {alt}) and downloading is unavailable offline. "
            f"Pass state_dict= or provide a valid ckpt_dir.")

    def _load_model(self):
        sd = self._resolve_state_dict()
        if not torch.cuda.is_available():                                      # fad.py:228-233, minus mps/cpu
            raise RuntimeError("frechet_audio_distance_exported_b200 needs a B200 GPU: there is no CPU fallback")
        self.engine = Engine(self.model_name, sd, precision=self.precision)
        self.model = _B200Model(self.engine)

    # ------------------------------------------------------------------ fad.py:302-408
    def _check_length(self, n: int) -> None:
        """length limits at the model's own sample rate (after any resampling)"""
        if self.model_name != "vggish" and self.model_name != "clap":
            n_fft = {8000: 256, 16000: 512, 32000: 1024}[self.sample_rate]
            if n <= n_fft // 2:
                raise ValueError("clip too short for reflect padding")

    def _prepare_clip(self, audio: np.ndarray, sr: int, device_resample: bool = False) -> np.ndarray:
        """Host part of vggish.py:241-250 / pann.py:93-101: mono mix, then resampling to the model's rate — on the
        host, or left to the GPU (`device_resample`: the clip comes back at its own rate, float32)."""
        is_pcm16 = isinstance(audio, RawPCM16)
        audio = np.asarray(audio)
        raw16 = is_pcm16 and audio.ndim == 1 and sr == self.sample_rate
        if is_pcm16 and not raw16:                                             # raw PCM16 that needs host-side work first
            audio = audio / 32768.0                                            # fad.py:148-149
        if audio.ndim > 1:                                                     # vggish.py:245-246 / pann.py:96-97
            audio = np.mean(audio, axis=1)
        if sr != self.sample_rate and self.model_name == "clap":
            # the reference pads to 480000 samples AT THE SOURCE RATE (fad.py:355-359) and int16-truncates
            # (clap.py:70-72) before it resamples (clap.py:75-80); the front end's own quantisation is switched off for
            # these clips (_embed_group).  Only the first 1001 frames survive (_pad_to_clap_time, fad.py:87-89), so the
            # source is cut to what they can see plus a margin far wider than the resampling filter.
            if audio.shape[0] < _CLAP_MAX_SAMPLES:
                audio = np.pad(audio, (0, _CLAP_MAX_SAMPLES - audio.shape[0]))
            keep = int((_CLAP_MAX_SAMPLES + 4096) * (float(sr) / float(self.sample_rate))) + 4096
            audio = audio[:keep].astype(np.float32)
            audio = (audio * np.float32(32767.0)).astype(np.int16).astype(np.float32) / np.float32(32767.0)
        if sr != self.sample_rate:                                             # vggish.py:249-250 / pann.py:100-101
            if device_resample:
                n_out = int(audio.shape[0] * (float(self.sample_rate) / float(sr)))
                if n_out < 1:
                    raise ValueError(f"Input signal length={audio.shape[0]} is too small to resample from "
                                     f"{sr}->{self.sample_rate}")
                self._check_length(n_out)
                return np.ascontiguousarray(audio, dtype=np.float32)
            audio = resample(audio, sr, self.sample_rate)
        # mono native-rate int16 PCM goes to the device as is (half the bytes); the front end divides by 32768
        audio = np.ascontiguousarray(audio, dtype=np.int16 if raw16 else np.float32)
        self._check_length(audio.shape[0])
        return audio

    def _canonical_length(self, a: np.ndarray) -> np.ndarray:
        """Clips of different lengths that the model cannot tell apart get ONE length, so that a directory of ragged
        clips becomes a handful of device batches instead of one per distinct length.  VGGish keeps whole 0.96 s patches
        only (vggish.py:268-277): samples past the last complete patch are never read, so clips are cut to
        (96 P - 1) * 160 + 400 samples.  CLAP input is zero-padded to 480 000 samples anyway (fad.py:355-359).  Both are
        exact; CNN14/PANN pools over every frame of the clip and keeps its own length."""
        n = a.shape[0]
        if self.model_name == "vggish":
            patches = self.engine.frontend_rows(n)
            need = (96 * patches - 1) * 160 + 400
            if 0 < patches and need < n:
                return a[:need]
        elif self.model_name == "clap" and n < _CLAP_MAX_SAMPLES:
            return np.pad(a, (0, _CLAP_MAX_SAMPLES - n))
        return a

    def get_embeddings(self, x: List[np.ndarray], sr: int) -> np.ndarray:
        """Embeddings for a list of clips, concatenated in input order.  Clips of equal length are
        batched into one device call (the reference loops clip by clip, fad.py:317); clips at another sample
        rate are resampled on the GPU (`fadb_resample`, same arithmetic as resample.py)."""
        on_device = sr != self.sample_rate
        prepared = []
        for audio in x:
            try:
                prepared.append(self._prepare_clip(audio, sr, device_resample=on_device))
            except Exception as e:                                             # fad.py:400-403
                if self.verbose:
                    print(f"[Exported FAD] Error processing audio: {e}")
                prepared.append(None)
        if not on_device:
            prepared = [a if a is None else self._canonical_length(a) for a in prepared]
        by_len: Dict[int, List[int]] = {}
        for i, a in enumerate(prepared):
            if a is not None:
                by_len.setdefault((a.shape[0], a.dtype.str), []).append(i)
        results: Dict[int, np.ndarray] = {}
        for (n, _), idxs in by_len.items():
            try:
                n_model = int(n * (float(self.sample_rate) / float(sr))) if on_device else n
                rows = self.engine.frontend_rows(n_model) if self.model_name == "vggish" else 1
                if rows <= 0:
                    for i in idxs:
                        results[i] = np.zeros((0, self.engine.dim), dtype=np.float32)
                    continue
                emb = self._embed_group([prepared[i] for i in idxs], rows, sr if on_device else None)
                for j, i in enumerate(idxs):
                    results[i] = emb[j * rows:(j + 1) * rows]
            except Exception as e:
                if self.verbose:
                    print(f"[Exported FAD] Error processing audio: {e}")
        embd_lst = [results[i] for i in range(len(prepared)) if i in results]
        if not embd_lst:
            return np.array([])                                                # fad.py:405-406
        return np.concatenate(embd_lst, axis=0)                                # fad.py:408

    def _embed_group(self, clips: List[np.ndarray], rows: int, resample_from: Optional[int]) -> np.ndarray:
        """Equal-length clips -> [len(clips) * rows, d] embeddings, pipelined: the host gathers chunk i+1 into a pinned
        staging buffer while chunk i is copied (copy stream) and chunk i-1 is embedded; the embeddings come back through
        a pinned buffer with asynchronous copies and ONE synchronisation at the end (the reference does a blocking
        round trip per clip, fad.py:389-396)."""
        eng = self.engine
        n, length, dtype = len(clips), clips[0].shape[0], torch.from_numpy(clips[0][:1]).dtype
        chunk = max(1, min(n, self._chunk_clips(length, staged=True)))
        depth = 3
        # pinned staging: flat byte buffers that only grow, viewed as [chunk, length] of this group's dtype — pinning
        # costs about a millisecond per megabyte, and a ragged directory is many groups (one per length)
        need = chunk * length * clips[0].dtype.itemsize
        flat = getattr(self, "_stage_flat", None)
        if flat is None or flat[0].numel() < need:
            if flat is not None:
                for used, ev in zip(self._stage_used, self._stage_ev):
                    if used:
                        ev.synchronize()
            flat = self._stage_flat = [torch.empty(need, dtype=torch.uint8).pin_memory() for _ in range(depth)]
            self._stage_ev = [torch.cuda.Event() for _ in range(depth)]
            self._stage_used = [False] * depth
        st = [f[:need].view(dtype).view(chunk, length) for f in flat]
        out_flat = getattr(self, "_out_flat", None)
        if out_flat is None or out_flat.numel() < n * rows * eng.dim:
            out_flat = self._out_flat = torch.empty(max(n * rows * eng.dim, 1 << 20), dtype=torch.float32).pin_memory()
        out_host = out_flat[: n * rows * eng.dim].view(n * rows, eng.dim)
        ring = self._ring()
        cur = torch.cuda.current_stream()
        prequantised = resample_from is not None and self.model_name == "clap"      # see _prepare_clip
        if prequantised:
            eng.set_clap_quantize(False)
        try:
            self._embed_chunks(clips, n, chunk, rows, depth, st, ring, out_host, resample_from)
        finally:
            if prequantised:
                eng.set_clap_quantize(True)
        cur.synchronize()
        return out_host.numpy().copy()             # the pinned buffer is reused by the next call

    def _embed_chunks(self, clips, n, chunk, rows, depth, st, ring, out_host, resample_from) -> None:
        eng = self.engine
        slot = 0
        for c0 in range(0, n, chunk):
            nc = min(chunk, n - c0)
            if self._stage_used[slot]:
                self._stage_ev[slot].synchronize()                 # the copy that last read this pinned buffer is done
            _gather_clips(st[slot].numpy(), clips, c0, nc)

            def consume(dev, _c0, _nc, c0=c0):
                pcm = eng.resample(dev, resample_from, self.sample_rate) if resample_from else dev
                emb = eng.embed_pcm(pcm)
                out_host[c0 * rows:(c0 + _nc) * rows].copy_(emb, non_blocking=True)

            ring.run(st[slot][:nc], nc, consume)
            self._stage_ev[slot].record(ring.stream)
            self._stage_used[slot] = True
            slot = (slot + 1) % depth

    def _chunk_clips(self, n_samples: int, staged: bool = False) -> int:
        """clips per host->device chunk: about 320 MB of fp32 PCM (512 ten-second 16 kHz clips were measured best on a
        B200, 256 .. 1024 within 2 %).  `staged` (get_embeddings: the chunk passes through pinned staging buffers that
        this object allocates) halves it — pinning costs ~1 ms per MB, once."""
        cap, nbytes = (256, 160 << 20) if staged else (512, 320 << 20)
        return max(1, min(cap, nbytes // max(4 * n_samples, 1)))

    def _ring(self) -> HostRing:
        if getattr(self, "_host_ring", None) is None:
            self._host_ring = HostRing(self.device, depth=4)
        return self._host_ring

    def _get_embedding_for_audio(self, audio: np.ndarray) -> np.ndarray:
        """fad.py:410-481."""
        a = self._prepare_clip(audio, self.sample_rate)
        return self.engine.embed_pcm(torch.from_numpy(a)[None].to(self.device)).cpu().numpy()

    # ------------------------------------------------------------------ fad.py:483-496
    def calculate_embd_statistics(self, embd_lst):
        if isinstance(embd_lst, list):
            embd_lst = np.array(embd_lst)
        embd_lst = np.asarray(embd_lst)
        in_dtype = embd_lst.dtype
        work = torch.float64 if in_dtype == np.float64 else torch.float32
        e = torch.from_numpy(np.ascontiguousarray(embd_lst)).to(self.device, work)
        eng = self.engine
        d = e.shape[1]
        acc = eng.new_acc(d)
        eng.stats_accumulate(e, acc)
        mu, sigma = eng.stats_finalize(acc, d)
        mu_np = mu.cpu().numpy()
        if in_dtype != np.float64:
            mu_np = mu_np.astype(np.float32)        # np.mean of float32 embeddings is float32
        return mu_np, sigma.cpu().numpy()

    # ------------------------------------------------------------------ fad.py:498-555
    def calculate_frechet_distance(self, mu1, sigma1, mu2, sigma2, eps: float = 1e-6) -> float:
        mu1 = np.atleast_1d(mu1)
        mu2 = np.atleast_1d(mu2)
        sigma1 = np.atleast_2d(sigma1)
        sigma2 = np.atleast_2d(sigma2)
        assert mu1.shape == mu2.shape, "Training and test mean vectors have different lengths"
        assert sigma1.shape == sigma2.shape, "Training and test covariances have different dimensions"
        dev = self.device
        t = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev) for a in (mu1, sigma1, mu2, sigma2)]
        out = self.engine.frechet(*t).cpu().numpy()
        if not np.isfinite(out[0]):
            raise ValueError("Frechet distance is not finite")
        return np.float64(out[0])

    # ------------------------------------------------------------------ fad.py:557-591
    def _load_audio_files(self, dir: str, dtype: str = "float32") -> List[np.ndarray]:
        pool = ThreadPool(self.audio_load_worker)
        files = [f for f in os.listdir(dir) if not f.startswith(".")]
        if self.verbose:
            print(f"[Exported FAD] Loading audio from {dir}...")
        tasks = [pool.apply_async(load_audio, args=(os.path.join(dir, f), self.sample_rate, self.channels, dtype,
                                                    dtype == "int16")) for f in files]
        pool.close()
        pool.join()
        return [t.get() for t in tasks]

    # ------------------------------------------------------------------ fad.py:593-662
    def score(self, background_dir: str, eval_dir: str, background_embds_path: Optional[str] = None,
              eval_embds_path: Optional[str] = None, dtype: str = "float32") -> float:
        try:
            if background_embds_path and os.path.exists(background_embds_path):
                if self.verbose:
                    print(f"[Exported FAD] Loading embeddings from {background_embds_path}...")
                embds_background = np.load(background_embds_path)
            else:
                audio_background = self._load_audio_files(background_dir, dtype=dtype)
                embds_background = self.get_embeddings(audio_background, sr=self.sample_rate)
                if background_embds_path:
                    os.makedirs(os.path.dirname(background_embds_path), exist_ok=True)
                    np.save(background_embds_path, embds_background)
            if eval_embds_path and os.path.exists(eval_embds_path):
                if self.verbose:
                    print(f"[Exported FAD] Loading embeddings from {eval_embds_path}...")
                embds_eval = np.load(eval_embds_path)
            else:
                audio_eval = self._load_audio_files(eval_dir, dtype=dtype)
                embds_eval = self.get_embeddings(audio_eval, sr=self.sample_rate)
                if eval_embds_path:
                    os.makedirs(os.path.dirname(eval_embds_path), exist_ok=True)
                    np.save(eval_embds_path, embds_eval)
            if len(embds_background) == 0:
                print("[Exported FAD] Background set dir is empty, exiting...")
                return -1
            if len(embds_eval) == 0:
                print("[Exported FAD] Eval set dir is empty, exiting...")
                return -1
            mu_background, sigma_background = self.calculate_embd_statistics(embds_background)
            mu_eval, sigma_eval = self.calculate_embd_statistics(embds_eval)
            return self.calculate_frechet_distance(mu_background, sigma_background, mu_eval, sigma_eval)
        except Exception as e:
            print(f"[Exported FAD] An error occurred: {e}")
            return -1

    # ------------------------------------------------------------------ B200 extensions
    def accumulate_clips(self, clips: torch.Tensor, acc: torch.Tensor, chunk_clips: Optional[int] = None) -> None:
        """Embed `clips` ([n, samples] fp32 or raw int16 PCM, HOST or device) and add their rows to the fp64 statistics
        buffer `acc`.  Host tensors are streamed in chunks through a ring of four device buffers on a copy
        stream, so the host->device copies run ahead of the kernels (pin the host tensor for this to be
        asynchronous).  Embeddings never leave the GPU."""
        eng = self.engine
        n = clips.shape[0]
        if n == 0:
            return
        if clips.is_cuda:
            eng.stats_accumulate(eng.embed_pcm(clips if clips.dtype == torch.int16 else clips.to(torch.float32)), acc)
            return
        assert clips.dtype in (torch.float32, torch.int16) and clips.dim() == 2 and clips.is_contiguous()
        if chunk_clips is None:
            chunk_clips = self._chunk_clips(clips.shape[1])
        chunk = min(chunk_clips, n)
        # ring of 4 device buffers on one copy stream (stream.HostRing): the copy engine runs up to four chunks ahead
        # and only ever waits for the kernels that last READ the buffer it is about to overwrite; events persist across
        # calls, so the first copy of a set overlaps the tail of the previous set.  The first chunk of a call is half
        # size: the kernels start after half a chunk's copy time.
        self._ring().run(clips, chunk, lambda dev, c0, nc: eng.stats_accumulate(eng.embed_pcm(dev), acc),
                         first=max(1, chunk // 2))

    def score_clips(self, background: torch.Tensor, evalset: torch.Tensor, reduce: bool = True) -> float:
        """FAD of two in-memory clip sets ([n, samples] fp32 or raw int16 PCM, host or device).  With torch.distributed
        initialised each rank passes ITS shard (see dist.shard_bounds): it embeds the shard,
        accumulates {n, sum x, sum x x^T} in fp64, ONE all-reduce over `process_group`, then every
        rank finalises mean/covariance and the Frechet distance redundantly.  reduce=False skips the all-reduce
        (the rank scores what it was given, alone)."""
        eng = self.engine
        d = eng.dim
        both = torch.zeros(2 * (1 + d + d * d), dtype=torch.float64, device=self.device)
        half = both.numel() // 2
        self.accumulate_clips(background, both[:half])
        self.accumulate_clips(evalset, both[half:])
        if reduce:
            eng.allreduce_acc(both, self.process_group)
        mu1, s1 = eng.stats_finalize(both[:half], d)
        mu2, s2 = eng.stats_finalize(both[half:], d)
        return float(eng.frechet(mu1, s1, mu2, s2)[0].item())
