"""Deterministic synthetic weights / inputs for the CFM/DiT path (TEST INFRASTRUCTURE ONLY).

``estimator_param_shapes`` restates the reference's parameter inventory
(models/estimator.py:66-96, models/diffusion_transformer.py:20-21,43-51,92-96;
SURVEY.md §8a: 116 tensors).  ``make_state`` fills them from a seeded CPU
generator with the reference's default-init scale (U(±1/sqrt(fan_in))) and —
because ``Decoder.initialize_weights`` zeroes the adaLN gates
(models/estimator.py:98-101), which would make attention/FFN weights
unobservable — draws ``adaLN_modulation.2`` from N(0, 0.3²).

There is no network and checkpoints are 80 MB, so tests regenerate weights from
the seed on each box; fixtures carry a checksum to detect RNG drift.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import torch


def estimator_param_shapes(n_mel: int = 80, hidden: int = 256, filt: int = 1024, n_layers: int = 6,
                           kernel: int = 3, gin: int = 256) -> "OrderedDict[str, Tuple[int, ...]]":
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def wb(name, *shape):
        s[name + ".weight"] = tuple(shape)
        s[name + ".bias"] = (shape[0],)

    wb("time_mlp.layer.0", filt, hidden)
    wb("time_mlp.layer.2", hidden, filt)
    wb("in_proj", hidden, hidden + n_mel, 1)
    for i in range(n_layers):
        p = f"blocks.{i}."
        wb(p + "time_fusion.film", 2 * hidden, hidden, 1)
        for n in "qkvo":
            wb(p + f"block.attn.conv_{n}", hidden, hidden, 1)
        wb(p + "block.mlp.conv_1", filt, hidden, kernel)
        wb(p + "block.mlp.conv_2", hidden, filt, kernel)
        if gin != hidden:
            wb(p + "block.adaLN_modulation.0", hidden, gin)
        wb(p + "block.adaLN_modulation.2", 6 * hidden, hidden)
    wb("final_proj", n_mel, hidden, 1)
    wb("cond_proj.0", filt, n_mel, kernel)
    wb("cond_proj.2", filt, filt, kernel)
    wb("cond_proj.4", hidden, filt, kernel)
    for i in range(n_layers // 2):
        wb(f"lsc_layers.{i}", hidden, 2 * hidden, kernel)
    return s


def make_state(seed: int = 0, n_mel: int = 80, dtype=torch.float32, adaln_std: float = 0.3,
               **dims) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    shapes = estimator_param_shapes(n_mel=n_mel, **dims)
    state: Dict[str, torch.Tensor] = OrderedDict()
    fan_in = {}
    for name, shape in shapes.items():
        base = name.rsplit(".", 1)[0]
        if name.endswith(".weight"):
            fi = 1
            for d in shape[1:]:
                fi *= d
            fan_in[base] = fi
        bound = 1.0 / (fan_in[base] ** 0.5)
        if "adaLN_modulation.2" in name:
            t = torch.randn(shape, generator=g) * adaln_std
        else:
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        state[name] = t.to(dtype)
    return state


def make_cfg_params(seed: int, n_mel: int = 80, gin: int = 256):
    """fake_speaker (1,gin), fake_content (1,n_mel,1) ~ N(0,1) (models/model.py:43-44 are
    zeros at init but learned; SURVEY.md §8d uses N(0,1))."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, gin, generator=g), torch.randn(1, n_mel, 1, generator=g)


def make_inputs(seed: int, lengths, T: int, n_mel: int = 80, gin: int = 256, t_per_sample: bool = False,
                t_value: float = 0.37):
    """Inputs as SURVEY.md §8d: mu ~ N(0,1) zeroed beyond len, c ~ N(0,1), x (=z) ~ N(0,1)
    UNMASKED (noise in the padded region, models/flow_matching.py:45), prefix mask."""
    g = torch.Generator().manual_seed(seed)
    B = len(lengths)
    lens = torch.as_tensor(list(lengths), dtype=torch.long)
    mask = (torch.arange(T)[None, :] < lens[:, None]).float().unsqueeze(1)     # utils/mask.py:4-8
    mu = torch.randn(B, n_mel, T, generator=g) * mask
    c = torch.randn(B, gin, generator=g)
    x = torch.randn(B, n_mel, T, generator=g)
    if t_per_sample:
        t = torch.rand(B, generator=g)
    else:
        t = torch.tensor(t_value)
    return dict(x=x, mu=mu, c=c, mask=mask, t=t, lengths=lens)


def checksum(state_or_tensors) -> float:
    vals = state_or_tensors.values() if isinstance(state_or_tensors, dict) else state_or_tensors
    acc = 0.0
    for i, v in enumerate(vals):
        acc += float(v.double().sum()) * (1.0 + 1e-3 * (i % 7)) + float(v.double().abs().sum()) * 1e-3
    return acc
