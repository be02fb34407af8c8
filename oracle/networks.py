"""Oracle (CPU, torch fp32) restatement of the reference embedding networks.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

Works on plain state_dicts whose keys are exactly those of the reference modules, so the same
dict can be (a) loaded into the reference's own nn.Modules (oracle/make_golden.py does this to pin
this file), (b) run here, (c) handed to the B200 weight packer.

Reference sites (relative to /root/reference/frechet_audio_distance_exported):
  * VGGishCore: models/vggish.py:40-51 (_make_layers), :54-95 (module + NHWC flatten).
  * ConvBlock / PANNCore (CNN14): models/pann.py:152-193, :200-273.
  * CLAP CNN14 head: README.md:195-199 of the reference (no code exists) — PARITY UNPINNED.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

VGGISH_CFG = [64, "M", 128, "M", 256, 256, "M", 512, 512, "M"]      # vggish.py:44
VGGISH_CONV_IDX = [0, 3, 6, 8, 11, 13]                               # nn.Sequential slots of the convs
CNN14_CHANNELS = [1, 64, 128, 256, 512, 1024, 2048]                  # pann.py:226-231
BN_EPS = 1e-5


# ----------------------------------------------------------------------------------------------
# Seeded, variance-preserving random weights (no checkpoints are obtainable offline)
# ----------------------------------------------------------------------------------------------
def _he(gen, shape, fan_in):
    return torch.randn(shape, generator=gen, dtype=torch.float32) * math.sqrt(2.0 / fan_in)


def vggish_random_state_dict(seed: int = 0, bias_std: float = 0.05) -> Dict[str, torch.Tensor]:
    """state_dict with the key names of VGGishCore (vggish.py:69-78)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    cin = 1
    for slot, cout in zip(VGGISH_CONV_IDX, [c for c in VGGISH_CFG if c != "M"]):
        sd[f"features.{slot}.weight"] = _he(g, (cout, cin, 3, 3), cin * 9)
        sd[f"features.{slot}.bias"] = torch.randn(cout, generator=g) * bias_std
        cin = cout
    for slot, (fin, fout) in zip([0, 2, 4], [(12288, 4096), (4096, 4096), (4096, 128)]):
        sd[f"embeddings.{slot}.weight"] = _he(g, (fout, fin), fin)
        sd[f"embeddings.{slot}.bias"] = torch.randn(fout, generator=g) * bias_std
    return sd


def _bn_random(g, prefix, c, sd):
    sd[f"{prefix}.weight"] = torch.rand(c, generator=g) + 0.5
    sd[f"{prefix}.bias"] = torch.randn(c, generator=g) * 0.1
    sd[f"{prefix}.running_mean"] = torch.randn(c, generator=g) * 0.1
    sd[f"{prefix}.running_var"] = torch.rand(c, generator=g) + 0.5
    sd[f"{prefix}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def cnn14_random_state_dict(seed: int = 0, bias_std: float = 0.05, clap_head: bool = False):
    """state_dict with the key names of PANNCore (pann.py:223-234); optional CLAP head keys
    `clap_head.0.*`, `clap_head.2.*` (our naming — the reference has no such module)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    # bn0 acts on dB log-mel values (~[-100, 40]); centre it so the net sees O(1) inputs.
    sd["bn0.weight"] = torch.rand(64, generator=g) + 0.5
    sd["bn0.bias"] = torch.randn(64, generator=g) * 0.1
    sd["bn0.running_mean"] = -30.0 + torch.randn(64, generator=g) * 5.0
    sd["bn0.running_var"] = 200.0 + torch.rand(64, generator=g) * 200.0
    sd["bn0.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    for b in range(1, 7):
        cin, cout = CNN14_CHANNELS[b - 1], CNN14_CHANNELS[b]
        sd[f"conv_block{b}.conv1.weight"] = _he(g, (cout, cin, 3, 3), cin * 9)
        sd[f"conv_block{b}.conv2.weight"] = _he(g, (cout, cout, 3, 3), cout * 9)
        _bn_random(g, f"conv_block{b}.bn1", cout, sd)
        _bn_random(g, f"conv_block{b}.bn2", cout, sd)
    sd["fc1.weight"] = _he(g, (2048, 2048), 2048)
    sd["fc1.bias"] = torch.randn(2048, generator=g) * bias_std
    if clap_head:
        sd["clap_head.0.weight"] = _he(g, (512, 2048), 2048)
        sd["clap_head.0.bias"] = torch.randn(512, generator=g) * bias_std
        sd["clap_head.2.weight"] = _he(g, (512, 512), 512)
        sd["clap_head.2.bias"] = torch.randn(512, generator=g) * bias_std
    return sd


# ----------------------------------------------------------------------------------------------
# Forward passes
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def vggish_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, return_intermediates: bool = False):
    """x [B,1,96,64] fp32 -> [B,128] fp32 — vggish.py:80-95."""
    acts = {}
    h = x
    slot = 0
    for v in VGGISH_CFG:
        if v == "M":
            h = F.max_pool2d(h, kernel_size=2, stride=2)
            acts[f"pool{slot}"] = h
            slot += 1
        else:
            h = F.relu(F.conv2d(h, sd[f"features.{slot}.weight"], sd[f"features.{slot}.bias"], padding=1))
            acts[f"conv{slot}"] = h
            slot += 2
    h = h.permute(0, 2, 3, 1).contiguous().view(h.shape[0], -1)      # NHWC flatten, vggish.py:91-94
    acts["flat"] = h
    h = F.relu(F.linear(h, sd["embeddings.0.weight"], sd["embeddings.0.bias"]))
    acts["fc1"] = h
    h = F.relu(F.linear(h, sd["embeddings.2.weight"], sd["embeddings.2.bias"]))
    acts["fc2"] = h
    h = F.linear(h, sd["embeddings.4.weight"], sd["embeddings.4.bias"])   # no final ReLU, vggish.py:76-77
    if return_intermediates:
        return h, acts
    return h


def _bn_eval(x, sd, prefix):
    return F.batch_norm(x, sd[f"{prefix}.running_mean"], sd[f"{prefix}.running_var"],
                        sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], training=False, eps=BN_EPS)


@torch.no_grad()
def cnn14_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, return_intermediates: bool = False):
    """x [B,1,T,64] fp32 -> [B,2048] fp32 — pann.py:236-273 (eval-mode BN)."""
    acts = {}
    h = x.transpose(1, 3)
    h = _bn_eval(h, sd, "bn0")                                        # pann.py:249-251
    h = h.transpose(1, 3)
    acts["bn0"] = h
    for b in range(1, 7):
        p = f"conv_block{b}"
        h = F.relu(_bn_eval(F.conv2d(h, sd[f"{p}.conv1.weight"], None, padding=1), sd, f"{p}.bn1"))
        acts[f"b{b}c1"] = h
        h = F.relu(_bn_eval(F.conv2d(h, sd[f"{p}.conv2.weight"], None, padding=1), sd, f"{p}.bn2"))
        acts[f"b{b}c2"] = h
        if b < 6:                                                     # pann.py:255-260
            h = F.avg_pool2d(h, kernel_size=(2, 2))
        acts[f"b{b}"] = h
    h = h.mean(dim=3)                                                 # pann.py:263
    h = h.max(dim=2).values + h.mean(dim=2)                           # pann.py:266-268
    acts["pooled"] = h
    h = F.relu(F.linear(h, sd["fc1.weight"], sd["fc1.bias"]))         # pann.py:271
    if return_intermediates:
        return h, acts
    return h


@torch.no_grad()
def clap_cnn14_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor):
    """x [B,1,1001,64] -> [B,512] L2-normalised.  README.md:195-199 of the reference
    (Linear(2048,512) -> ReLU -> Linear(512,512) -> L2 norm on top of CNN14).  PARITY UNPINNED."""
    h = cnn14_forward(sd, x)
    h = F.relu(F.linear(h, sd["clap_head.0.weight"], sd["clap_head.0.bias"]))
    h = F.linear(h, sd["clap_head.2.weight"], sd["clap_head.2.bias"])
    return F.normalize(h, dim=-1)
