"""Drop-in for the reference's ``models.estimator.Decoder`` (models/estimator.py:65-137).

Same constructor signature, same ``forward(t, x, mask, mu, c)``, same parameter names/shapes
(``state_dict()`` carries the reference's 116 ``estimator.*``-relative keys, so checkpoints saved
by the reference load unchanged, api.py:49) — but the computation is one call into the sm_100a
CUDA library through the C ABI (include/stabletts_b200.h).  Parameters live in ordinary
``nn.Parameter`` s (so ``.to()``, ``.eval()``, ``.parameters()`` behave); the library keeps its own
packed copy which is refreshed whenever a parameter's version counter or device changes.

No CPU fallback: tensors must be CUDA fp32; anything else raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections import OrderedDict
from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import _lib


def _param_shapes(n_mel, hidden, filt, n_layers, kernel, gin) -> "OrderedDict[str, Tuple[int, ...]]":
    """Parameter inventory of Decoder.__init__ (models/estimator.py:66-96) in registration order."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def wb(name, *shape):
        s[name + ".weight"] = tuple(shape)
        s[name + ".bias"] = (shape[0],)

    wb("time_mlp.layer.0", filt, hidden)                       # :55-59
    wb("time_mlp.layer.2", hidden, filt)
    wb("in_proj", hidden, hidden + n_mel, 1)                   # :78
    for i in range(n_layers):                                  # :79
        p = f"blocks.{i}."
        wb(p + "time_fusion.film", 2 * hidden, hidden, 1)      # :28
        for n in "qkv":
            wb(p + f"block.attn.conv_{n}", hidden, hidden, 1)  # diffusion_transformer.py:43-45
        wb(p + "block.attn.conv_o", hidden, hidden, 1)         # :51
        wb(p + "block.mlp.conv_1", filt, hidden, kernel)       # :20
        wb(p + "block.mlp.conv_2", hidden, filt, kernel)       # :21
        wb(p + "block.adaLN_modulation.2", 6 * hidden, hidden)  # :92-96 (0 = Identity, 1 = SiLU)
    wb("final_proj", n_mel, hidden, 1)                         # :80
    wb("cond_proj.0", filt, n_mel, kernel)                     # :83-89
    wb("cond_proj.2", filt, filt, kernel)
    wb("cond_proj.4", hidden, filt, kernel)
    for i in range(n_layers // 2):
        wb(f"lsc_layers.{i}", hidden, 2 * hidden, kernel)      # :94
    return s


class _Node(nn.Module):
    """Anonymous container: exists only so parameter paths equal the reference's module tree."""


class Decoder(nn.Module):
    def __init__(self, noise_channels, cond_channels, hidden_channels, out_channels, filter_channels, dropout=0.1,
                 n_layers=1, n_heads=4, kernel_size=3, gin_channels=0, use_lsc=True):
        super().__init__()
        if not (noise_channels == cond_channels == out_channels):
            raise ValueError("noise/cond/out channels must all equal n_mel (as models/model.py:40 builds it)")
        if not use_lsc:
            raise ValueError("use_lsc=False is not built (the reference never constructs it)")
        if gin_channels != hidden_channels:
            raise ValueError("gin_channels must equal hidden_channels (adaLN_modulation.0 is Identity)")
        self.noise_channels = noise_channels
        self.cond_channels = cond_channels
        self.hidden_channels = hidden_channels
        self.out_channels = out_channels
        self.filter_channels = filter_channels
        self.use_lsc = use_lsc
        self.n_layers, self.n_heads, self.kernel_size, self.gin_channels = n_layers, n_heads, kernel_size, gin_channels
        self.n_lsc_layers = n_layers // 2
        self._shapes = _param_shapes(noise_channels, hidden_channels, filter_channels, n_layers, kernel_size, gin_channels)
        for name, shape in self._shapes.items():
            self._register(name, nn.Parameter(torch.empty(shape)))
        self.initialize_weights()
        # library state (not part of the module state)
        self._handle = None
        self._handle_device = None
        self._synced: Dict[str, Tuple[int, int]] = {}
        self._workspace = None
        # debugging switch between the two CUDA GEMM engines (not a backend dispatch; both are this library)
        self._engine = _lib.ST_ENGINE_SIMT if os.environ.get("STABLETTS_B200_ENGINE") == "simt" else _lib.ST_ENGINE_TCGEN05

    # -- module tree ------------------------------------------------------------------------------
    def _register(self, dotted: str, p: nn.Parameter) -> None:
        mod = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            if part not in mod._modules:
                mod.add_module(part, _Node())
            mod = mod._modules[part]
        mod.register_parameter(parts[-1], p)

    def _param(self, dotted: str) -> nn.Parameter:
        mod = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            mod = mod._modules[part]
        return mod._parameters[parts[-1]]

    def initialize_weights(self):
        """PyTorch default Conv1d/Linear init (U(±1/sqrt(fan_in)) for weight and bias), xavier on the
        q/k/v projections (diffusion_transformer.py:54-56), zero adaLN gates (estimator.py:98-101)."""
        with torch.no_grad():
            for name, shape in self._shapes.items():
                p = self._param(name)
                base = name.rsplit(".", 1)[0]
                wshape = self._shapes[base + ".weight"]
                fan_in = 1
                for d in wshape[1:]:
                    fan_in *= d
                if "adaLN_modulation.2" in name:
                    p.zero_()
                elif name.endswith(".weight") and any(k in name for k in ("conv_q", "conv_k", "conv_v")):
                    nn.init.xavier_uniform_(p)
                else:
                    bound = 1.0 / math.sqrt(fan_in)
                    p.uniform_(-bound, bound)

    # -- library plumbing -------------------------------------------------------------------------
    def set_engine(self, name: str) -> None:
        """'tcgen05' (default product path) or 'simt' (fp32 cross-check engine) — both CUDA."""
        self._engine = {"tcgen05": _lib.ST_ENGINE_TCGEN05, "simt": _lib.ST_ENGINE_SIMT}[name]
        if self._handle is not None:
            lib = _lib.load_library()
            _lib.check(lib, self._handle, lib.st_set_engine(self._handle, self._engine), "st_set_engine")

    def _ensure_handle(self, device: torch.device):
        lib = _lib.load_library()
        if device.type != "cuda":
            raise RuntimeError("stabletts_b200 runs on CUDA (B200) only: there is no CPU fallback")
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is not None and self._handle_device != index:
            self.release()
        if self._handle is None:
            dims = _lib.StDims(self.noise_channels, self.hidden_channels, self.filter_channels, self.n_heads,
                               self.n_layers, self.kernel_size, self.gin_channels)
            h = C.c_void_p()
            rc = lib.st_create(C.byref(dims), index, C.byref(h))
            _lib.check(lib, None, rc, "st_create")
            self._handle, self._handle_device = h, index
            self._synced.clear()
            _lib.check(lib, h, lib.st_set_engine(h, self._engine), "st_set_engine")
        return lib, self._handle

    def _sync_weights(self, lib, h, stream: int) -> None:
        dirty = False
        for name in self._shapes:
            p = self._param(name)
            if p.device.type != "cuda" or p.dtype != torch.float32:
                raise RuntimeError(f"parameter {name} must be CUDA fp32 (got {p.device}, {p.dtype}); call .to('cuda')")
            tag = (p.data_ptr(), p._version)
            if self._synced.get(name) != tag:
                pc = p.detach().contiguous()
                _lib.check(lib, h, lib.st_load_weight(h, name.encode(), pc.data_ptr(), pc.numel(), stream),
                           f"st_load_weight({name})")
                self._synced[name] = tag
                dirty = True
        if dirty:
            _lib.check(lib, h, lib.st_finalize_weights(h, stream), "st_finalize_weights")

    def _ensure_workspace(self, lib, h, B: int, T: int, cfg: int, device) -> None:
        need = lib.st_workspace_bytes(h, B, T, cfg)
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != device:
            self._workspace = None
            self._workspace = torch.empty(need, dtype=torch.uint8, device=device)
            _lib.check(lib, h, lib.st_attach_workspace(h, self._workspace.data_ptr(), self._workspace.numel()),
                       "st_attach_workspace")

    def _prepare(self, ref: torch.Tensor, B: int, T: int, cfg: int):
        lib, h = self._ensure_handle(ref.device)
        stream = torch.cuda.current_stream(ref.device).cuda_stream
        self._sync_weights(lib, h, stream)
        self._ensure_workspace(lib, h, B, T, cfg, ref.device)
        return lib, h, stream

    def release(self) -> None:
        if self._handle is not None:
            _lib.load_library().st_destroy(self._handle)
        self._handle = None
        self._workspace = None
        self._synced.clear()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def launch_count(self) -> int:
        return 0 if self._handle is None else int(_lib.load_library().st_launch_count(self._handle))

    @staticmethod
    def _f32c(name: str, t: torch.Tensor, shape) -> torch.Tensor:
        if not isinstance(t, torch.Tensor) or t.device.type != "cuda":
            raise RuntimeError(f"{name} must be a CUDA tensor (no CPU fallback)")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
        return t.detach().to(torch.float32).contiguous()

    # -- the reference's forward ------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, t, x, mask, mu, c):
        """models/estimator.py:103-137.  t: 0-dim or (B,); x, mu: (B, n_mel, T); mask: (B, 1, T)
        float {0,1}; c: (B, gin).  Returns (B, n_mel, T), exactly 0 at masked frames."""
        B, M, T = x.shape
        x = self._f32c("x", x, (B, self.noise_channels, T))
        mu = self._f32c("mu", mu, (B, self.cond_channels, T))
        mask = self._f32c("mask", mask, (B, 1, T))
        c = self._f32c("c", c, (B, self.gin_channels))
        t = torch.as_tensor(t, dtype=torch.float32, device=x.device).reshape(-1).contiguous()
        if t.numel() not in (1, B):
            raise ValueError("t must be 0-dim or have shape (B,)")
        lib, h, stream = self._prepare(x, B, T, 0)
        out = torch.empty_like(x)
        rc = lib.st_estimator_forward(h, t.data_ptr(), t.numel(), x.data_ptr(), mask.data_ptr(), mu.data_ptr(),
                                      c.data_ptr(), out.data_ptr(), B, T, stream)
        _lib.check(lib, h, rc, "st_estimator_forward")
        return out
