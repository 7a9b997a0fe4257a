"""CPU oracle for SURVEY.md §8 row f2 (TEST INFRASTRUCTURE ONLY): functional restatement of
models/text_encoder.py:34-44 on top of the DiT block restatement in estimator_ref.py.  Pinned by
tests/test_text_encoder.py (live reference module + reference-generated fixtures)."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from oracle.estimator_ref import dit_block


def param_shapes(n_vocab=401, out_channels=80, hidden=256, filt=1024, n_layers=3, kernel=3):
    s = OrderedDict()

    def wb(name, *shape):
        s[name + ".weight"] = tuple(shape)
        s[name + ".bias"] = (shape[0],)

    s["emb.weight"] = (n_vocab, hidden)
    for i in range(n_layers):
        p = f"encoder.{i}."
        for n in "qkv":
            wb(p + f"attn.conv_{n}", hidden, hidden, 1)
        wb(p + "attn.conv_o", hidden, hidden, 1)
        wb(p + "mlp.conv_1", filt, hidden, kernel)
        wb(p + "mlp.conv_2", hidden, filt, kernel)
        wb(p + "adaLN_modulation.2", 6 * hidden, hidden)
    wb("proj", out_channels, hidden, 1)
    return s


def make_state(seed=3, adaln_std=0.3, **dims):
    g = torch.Generator().manual_seed(seed)
    shapes = param_shapes(**dims)
    st = OrderedDict()
    fan = {}
    for name, shape in shapes.items():
        if name == "emb.weight":
            st[name] = torch.randn(shape, generator=g) * shape[1] ** -0.5
            continue
        base = name.rsplit(".", 1)[0]
        if name.endswith(".weight"):
            fi = 1
            for d in shape[1:]:
                fi *= d
            fan[base] = fi
        if "adaLN_modulation.2" in name:
            st[name] = torch.randn(shape, generator=g) * adaln_std        # un-zero the gates (estimator.py-style fact 2)
        else:
            st[name] = (torch.rand(shape, generator=g) * 2 - 1) / fan[base] ** 0.5
    return st


def text_encoder_forward(state, ids, c, x_lengths, n_heads=4):
    """models/text_encoder.py:34-44."""
    hidden = state["emb.weight"].shape[1]
    n_layers = 1 + max(int(k.split(".")[1]) for k in state if k.startswith("encoder."))
    x = F.embedding(ids, state["emb.weight"]) * hidden ** 0.5           # :35
    x = x.transpose(1, -1)                                              # :36
    T = x.size(2)
    x_mask = (torch.arange(T, device=ids.device)[None, :] < x_lengths[:, None]).unsqueeze(1).to(x.dtype)   # :37
    for i in range(n_layers):
        x = dit_block(state, f"encoder.{i}.", x, c, x_mask, n_heads)    # :39-40
    mu_x = F.conv1d(x, state["proj.weight"], state["proj.bias"]) * x_mask   # :42
    return x, mu_x, x_mask


def make_inputs(seed, lens, T, n_vocab=401, gin=256):
    g = torch.Generator().manual_seed(seed)
    B = len(lens)
    ids = torch.randint(1, n_vocab, (B, T), generator=g)
    lens_t = torch.as_tensor(lens)
    ids = ids * (torch.arange(T)[None] < lens_t[:, None])              # padding id 0 beyond the length
    c = torch.randn(B, gin, generator=g)
    return ids, c, lens_t


CASES = {
    "tenc_basic":  dict(seed=51, lens=[129, 77], T=129, out_channels=80),
    "tenc_mel128": dict(seed=52, lens=[40, 40, 13], T=40, out_channels=128),
    "tenc_T1":     dict(seed=53, lens=[1], T=1, out_channels=80),
}
