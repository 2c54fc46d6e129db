"""Position embeddings of the coarse points that feed the transformer on CUDA kernels.

Same class names, constructor arguments and ``forward(xyz)`` contract as the reference's
``models/transformer/position_embedding.py`` (``PositionEmbeddingCoordsSine`` :8-49, ``PositionEmbeddingLearned``
:52-70, whose ``mlp.{0,2,4,6,8}`` parameter names are kept so checkpoints load), but the work is done by
``kpreg_sine_embed`` (one pass, no [N, 3, F] temporaries) and — for the learned variant in inference — by the tcgen05
Linear kernel with bias + ReLU in its epilogue.  ``utils/seq_manipulation.pack_coarse_level`` fuses the sine code with
the padding of the coarse level.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops


def sine_frequencies(num_pos_feats: int, temperature: float, device) -> torch.Tensor:
    """dim_t[k] = temperature^(2 (k // 2) / F), evaluated in fp32 by torch exactly as the reference evaluates it
    (position_embedding.py:38-39), so that the kernel's sin / cos arguments are bit-identical to the reference's."""
    k = torch.arange(num_pos_feats, dtype=torch.float32, device=device)
    return temperature ** (2 * torch.div(k, 2, rounding_mode='trunc') / num_pos_feats)


class PositionEmbeddingCoordsSine(nn.Module):
    """Continuous-coordinate generalisation of the transformer's sine code: channel d*F + k of the output is
    sin (k even) or cos (k odd) of ``xyz[..., d] * scale * 2 pi / temperature^(2 (k // 2) / F)`` with
    F = d_model // n_dim rounded down to an even number; the remaining ``d_model - n_dim * F`` channels are zero."""

    def __init__(self, n_dim: int = 1, d_model: int = 256, temperature=10000, scale=None):
        super().__init__()
        self.n_dim = n_dim
        self.d_model = d_model
        self.num_pos_feats = d_model // n_dim // 2 * 2
        self.temperature = temperature
        self.padding = d_model - self.num_pos_feats * self.n_dim
        self.scale = (1.0 if scale is None else scale) * 2 * math.pi
        self._dim_t = None

    def frequencies(self, device) -> torch.Tensor:
        if self._dim_t is None or self._dim_t.device != device:
            self._dim_t = sine_frequencies(self.num_pos_feats, self.temperature, device)
        return self._dim_t

    def forward(self, xyz: torch.Tensor) -> torch.Tensor:
        assert xyz.shape[-1] == self.n_dim
        return ops.sine_embed(xyz, self.d_model, self.num_pos_feats, self.scale, self.frequencies(xyz.device))


class PositionEmbeddingLearned(nn.Module):
    """Absolute position embedding from a 5-layer MLP (n_dim -> 32 -> 64 -> 128 -> 256 -> d_model, ReLU between)."""

    _WIDTHS = (32, 64, 128, 256)

    def __init__(self, n_dim: int = 1, d_model: int = 256):
        super().__init__()
        layers, prev = [], n_dim
        for width in self._WIDTHS:
            layers += [nn.Linear(prev, width), nn.ReLU()]
            prev = width
        layers.append(nn.Linear(prev, d_model))
        self.mlp = nn.Sequential(*layers)

    def forward(self, xyz: torch.Tensor) -> torch.Tensor:
        if not xyz.is_cuda or torch.is_grad_enabled():
            return self.mlp(xyz)
        x = xyz.reshape(-1, xyz.shape[-1]).to(torch.float32)
        linears = [m for m in self.mlp if isinstance(m, nn.Linear)]
        for i, lin in enumerate(linears):
            # the first layer's K (= n_dim, 3) is below the TMA path's minimum: kpreg_linear_forward falls back to its fp32 kernel
            x = ops.linear_forward(x, lin.weight, None, lin.bias, act="relu" if i + 1 < len(linears) else None)
        return x.reshape(*xyz.shape[:-1], x.shape[-1])
