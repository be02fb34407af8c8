"""Oracle end-to-end path: the reference's `get_embeddings` -> statistics -> Frechet chain on CPU.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  Also the CPU baseline that bench.py times
(`cpu_baseline`, `--impl reference`); it is never on the product path.

Restates fad.py:302-408 (get_embeddings: ONE CLIP PER ITERATION, batch = that clip's patches, as
the reference does — this is deliberately not batched across clips so the CPU baseline has the
reference's performance shape), fad.py:483-496, fad.py:498-555, fad.py:640-656.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from . import frontend, networks, stats


class OracleFAD:
    """Mirror of the reference `FrechetAudioDistance` hot-path methods on CPU.

    model_name: "vggish" | "pann-8k" | "pann-16k" | "pann-32k" | "clap" (CNN14 branch, README variant).
    """

    SR = {"vggish": 16000, "pann-8k": 8000, "pann-16k": 16000, "pann-32k": 32000, "clap": 48000}

    def __init__(self, model_name: str, state_dict: Dict[str, torch.Tensor]):
        if model_name not in self.SR:
            raise ValueError(f"Unknown model: {model_name}")
        self.model_name = model_name
        self.sample_rate = self.SR[model_name]
        self.sd = {k: v.detach().to(torch.float32) if v.is_floating_point() else v
                   for k, v in state_dict.items()}

    # fad.py:317-403, one clip
    def embed_clip(self, audio: np.ndarray) -> np.ndarray:
        if self.model_name == "vggish":
            ex = frontend.vggish_examples(audio)                              # fad.py:388
            x = torch.from_numpy(ex)[:, None, :, :]
            return networks.vggish_forward(self.sd, x).numpy()               # fad.py:392-396
        if self.model_name == "clap":
            feats = frontend.clap_features(audio)                            # fad.py:356-362
            x = torch.from_numpy(feats)[None, None]
            return networks.clap_cnn14_forward(self.sd, x).numpy()
        feats = frontend.pann_features(audio, self.sample_rate)              # fad.py:374-377
        x = torch.from_numpy(feats)[None, None]
        return networks.cnn14_forward(self.sd, x).numpy()                    # fad.py:381-385

    def get_embeddings(self, x: List[np.ndarray], sr: int = None) -> np.ndarray:
        out = []
        for audio in x:
            try:
                out.append(self.embed_clip(np.asarray(audio)))
            except Exception:                                                # fad.py:400-403
                continue
        if not out:
            return np.array([])                                              # fad.py:405-406
        return np.concatenate(out, axis=0)                                   # fad.py:408

    calculate_embd_statistics = staticmethod(stats.embd_statistics)
    calculate_frechet_distance = staticmethod(stats.frechet_distance)

    def fad_from_clips(self, background: List[np.ndarray], evalset: List[np.ndarray]):
        """fad.py:621-656 minus file loading; returns (fad, emb_bg, emb_ev)."""
        eb = self.get_embeddings(background)
        ee = self.get_embeddings(evalset)
        if len(eb) == 0 or len(ee) == 0:
            return -1, eb, ee
        mu1, s1 = stats.embd_statistics(eb)
        mu2, s2 = stats.embd_statistics(ee)
        return stats.frechet_distance(mu1, s1, mu2, s2), eb, ee
