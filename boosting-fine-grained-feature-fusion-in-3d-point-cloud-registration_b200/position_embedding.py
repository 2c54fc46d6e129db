"""Position embeddings of the coarse points fed to the transformer — mirror of the reference's
``models/transformer/position_embedding.py`` (:8-70): ``PositionEmbeddingCoordsSine`` (continuous-coordinate sine/cosine
code, no parameters) and ``PositionEmbeddingLearned`` (5-layer MLP, same sub-module names so checkpoints load)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class PositionEmbeddingCoordsSine(nn.Module):
    """pos_emb[..., (d * F + 2j, d * F + 2j + 1)] = (sin, cos)(xyz[..., d] * scale / T^(2j / F)), F = d_model // n_dim // 2 * 2,
    zero-padded to d_model (reference :8-49)."""

    def __init__(self, n_dim: int = 1, d_model: int = 256, temperature=10000, scale=None):
        super().__init__()
        self.n_dim = n_dim
        self.num_pos_feats = d_model // n_dim // 2 * 2
        self.temperature = temperature
        self.padding = d_model - self.num_pos_feats * self.n_dim
        if scale is None:
            scale = 1.0
        self.scale = scale * 2 * math.pi

    def forward(self, xyz: torch.Tensor) -> torch.Tensor:
        assert xyz.shape[-1] == self.n_dim
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=xyz.device)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode='trunc') / self.num_pos_feats)
        pos_divided = (xyz * self.scale).unsqueeze(-1) / dim_t
        pos_sin = pos_divided[..., 0::2].sin()
        pos_cos = pos_divided[..., 1::2].cos()
        pos_emb = torch.stack([pos_sin, pos_cos], dim=-1).reshape(*xyz.shape[:-1], -1)
        return F.pad(pos_emb, (0, self.padding))


class PositionEmbeddingLearned(nn.Module):
    """Absolute position embedding, learned (reference :52-70)."""

    def __init__(self, n_dim: int = 1, d_model: int = 256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(n_dim, 32), nn.ReLU(), nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 128), nn.ReLU(),
                                 nn.Linear(128, 256), nn.ReLU(), nn.Linear(256, d_model))

    def forward(self, xyz: torch.Tensor) -> torch.Tensor:
        return self.mlp(xyz)
