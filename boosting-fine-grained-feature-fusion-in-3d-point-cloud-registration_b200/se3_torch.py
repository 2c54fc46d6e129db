"""Rigid-transform estimation — mirror of the reference's ``utils/se3_torch.py`` pose solve.

``compute_rigid_transform(a, b, weights=None)`` (reference :131-173) and
``fast_compute_rigid_transform(a, b, weights=None, weights_threshold=0.85)`` (:226-274) keep their
signatures and semantics — including the fast variant's in-place zeroing of the caller's weights
(:240-242) — but all leading dimensions are solved by one launch of ``kpreg_kabsch`` (one CTA per
correspondence set) instead of a Python loop over pairs calling ``torch.svd``.
``compute_rigid_transform_batch`` solves ragged sets (different point counts per pair) in one launch.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

from . import ops

_EPS = 1e-6


def se3_init(rot=None, trans=None):
    assert rot is not None or trans is not None
    if rot is not None and trans is not None:
        return torch.cat([rot, trans], dim=-1)
    if rot is None:
        eye = torch.eye(3, device=trans.device, dtype=trans.dtype).expand(*trans.shape[:-2], 3, 3)
        return torch.cat([eye, trans], dim=-1)
    return torch.nn.functional.pad(rot, (0, 1))


def se3_cat(a, b):
    """Composition a∘b of [*,3,4] transforms."""
    rot = a[..., :3, :3] @ b[..., :3, :3]
    trans = a[..., :3, :3] @ b[..., :3, 3:4] + a[..., :3, 3:4]
    return torch.cat([rot, trans], dim=-1)


def se3_inv(pose):
    rot_t = pose[..., :3, :3].transpose(-1, -2)
    return torch.cat([rot_t, -rot_t @ pose[..., :3, 3:4]], dim=-1)


def se3_transform(pose, xyz):
    """Apply [*,3,4] pose to [*,N,3] points."""
    return xyz @ pose[..., :3, :3].transpose(-1, -2) + pose[..., :3, 3].unsqueeze(-2)


def se3_compare(a, b):
    """Rotation (degrees) and translation error of a∘b⁻¹ (reference :117-129)."""
    combined = se3_cat(a, se3_inv(b))
    trace = combined[..., 0, 0] + combined[..., 1, 1] + combined[..., 2, 2]
    rot_err_deg = torch.acos(torch.clamp(0.5 * (trace - 1), -1., 1.)) * 180 / math.pi
    trans_err = torch.norm(combined[..., :, 3], dim=-1)
    return {'rot_deg': rot_err_deg, 'trans': trans_err}


def _solve(a: torch.Tensor, b: torch.Tensor, weights: Optional[torch.Tensor], threshold: float, write_back: bool):
    assert a.shape == b.shape
    assert a.shape[-1] == 3
    if not a.is_cuda:
        raise RuntimeError("compute_rigid_transform: CUDA tensors required (kpreg_b200 has no CPU path)")
    lead = a.shape[:-2]
    n_pts = a.shape[-2]
    n_sets = 1
    for d in lead:
        n_sets *= int(d)
    w_flat = None
    if weights is not None:
        assert a.shape[:-1] == weights.shape
        w_flat = weights
        if write_back and not (weights.is_contiguous() and weights.dtype == torch.float32):
            # keep the reference's side effect even for odd layouts: threshold in torch, then solve
            weights.copy_(torch.where(weights > threshold, weights, torch.zeros_like(weights)))
            threshold, write_back = -1.0, False
        w_flat = weights.reshape(-1) if weights.is_contiguous() else weights.contiguous().reshape(-1)
    out = ops.kabsch(a.reshape(-1, 3), b.reshape(-1, 3), w_flat, n_sets, n_pts, None, threshold, write_back)
    return out.reshape(*lead, 3, 4).to(a.dtype)


def compute_rigid_transform(a: torch.Tensor, b: torch.Tensor, weights: torch.Tensor = None):
    """Transform T ([*,] 3, 4) with T*a = b in the weighted least-squares sense."""
    if weights is not None:
        assert a.shape[:-1] == weights.shape
        # the reference asserts 0 <= w <= 1 (a host sync on GPU tensors); kept for identical error behaviour
        assert weights.min() >= 0 and weights.max() <= 1
    return _solve(a, b, weights, -1.0, False)


def fast_compute_rigid_transform(a: torch.Tensor, b: torch.Tensor, weights: torch.Tensor = None,
                                 weights_threshold=0.85):
    """As compute_rigid_transform, after zeroing (in place) every weight that is not above the threshold."""
    assert a.shape == b.shape
    assert a.shape[-1] == 3
    if weights is None:
        return _solve(a, b, None, -1.0, False)
    assert a.shape[:-1] == weights.shape
    assert weights.min() >= 0 and weights.max() <= 1
    return _solve(a, b, weights, float(weights_threshold), True)


def compute_rigid_transform_batch(a: Sequence[torch.Tensor], b: Sequence[torch.Tensor],
                                  weights: Optional[Sequence[torch.Tensor]] = None,
                                  weights_threshold: Optional[float] = None) -> torch.Tensor:
    """Per-pair solve for a list of [L, N_p, 3] correspondence tensors with different N_p (RegTR's
    ``[fast_compute_rigid_transform(...) for b in range(B)]``, models/finegrained_regtr.py:215-218) in
    ONE launch.  Returns [L, B, 3, 4] like the reference's ``torch.stack(..., dim=1)``."""
    n_pairs = len(a)
    n_layers = int(a[0].shape[0])
    sizes = [int(t.shape[1]) for t in a]
    dev = a[0].device
    flat_a = torch.cat([t.reshape(-1, 3) for t in a], 0)
    flat_b = torch.cat([t.reshape(-1, 3) for t in b], 0)
    flat_w = torch.cat([t.reshape(-1) for t in weights], 0).to(torch.float32) if weights is not None else None
    offs = [0]
    for n in sizes:
        for _ in range(n_layers):
            offs.append(offs[-1] + n)
    offsets = torch.tensor(offs, dtype=torch.int64).to(dev, non_blocking=True)
    thr = -1.0 if weights_threshold is None or weights is None else float(weights_threshold)
    out = ops.kabsch(flat_a, flat_b, flat_w, n_pairs * n_layers, 0, offsets, thr, False)
    return out.reshape(n_pairs, n_layers, 3, 4).transpose(0, 1).contiguous()
