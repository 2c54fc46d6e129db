"""Getting clouds onto the device: the reference's ``.pth`` cloud format and its
``ShufflePoints`` transform, without a host round trip for the point data.

* ``load_cloud(path, device)`` — the 3DMatch / MCD loaders read a cloud with ``torch.load`` (a pickled numpy
  ``[N,3]`` array, reference data_loaders/threedmatch.py:74-75); here it is read once, pinned and copied
  asynchronously.
* ``ShufflePoints(max_pts, shuffle)`` — reference data_loaders/transforms.py:95-131: a random permutation truncated
  to ``max_pts`` applied to points and overlap masks of both clouds, correspondences remapped through the reverse
  indices and filtered.  The permutation is drawn on the host from numpy's global generator exactly as the reference
  does (same draws, same order: source first), only the 8-byte indices cross the bus; the gathers, the reverse index
  and the remap run in ``kpreg_shuffle_gather`` / ``kpreg_remap_pairs``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def load_cloud(path: str, device=None) -> torch.Tensor:
    """[N,3] float32 CUDA tensor from a reference-format ``.pth`` cloud (``torch.load`` of a numpy array or tensor)."""
    obj = torch.load(path, weights_only=False)
    cloud = torch.from_numpy(np.ascontiguousarray(obj, dtype=np.float32)) if isinstance(obj, np.ndarray) else obj.to(torch.float32)
    if cloud.dim() != 2 or cloud.shape[1] != 3:
        raise RuntimeError(f"{path}: expected an [N,3] cloud, got {tuple(cloud.shape)}")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    return cloud.contiguous().pin_memory().to(dev, non_blocking=True)


class ShufflePoints:
    """Shuffle the points (and drop all but ``max_pts``) of a pair dict in collate_pair's format, on the device."""

    def __init__(self, max_pts=30000, shuffle=True):
        self.max_pts = max_pts
        self.shuffle = shuffle

    def _indices(self, n: int) -> np.ndarray:
        if self.shuffle:
            return np.random.permutation(n)[:self.max_pts]
        return np.arange(min(n, self.max_pts))

    def __call__(self, data):
        dev = data['src_xyz'].device
        if dev.type != 'cuda':
            raise RuntimeError("ShufflePoints: clouds must be CUDA tensors (kpreg_b200 has no CPU path)")
        n_src, n_tgt = int(data['src_xyz'].shape[0]), int(data['tgt_xyz'].shape[0])
        src_idx, tgt_idx = self._indices(n_src), self._indices(n_tgt)          # the reference's draw order
        perm_s = torch.from_numpy(src_idx.astype(np.int64)).pin_memory().to(dev, non_blocking=True)
        perm_t = torch.from_numpy(tgt_idx.astype(np.int64)).pin_memory().to(dev, non_blocking=True)
        want_rev = 'correspondences' in data
        src, src_ov, rev_s, st_s = ops.shuffle_gather(data['src_xyz'], data.get('src_overlap'), perm_s, want_rev)
        tgt, tgt_ov, rev_t, st_t = ops.shuffle_gather(data['tgt_xyz'], data.get('tgt_overlap'), perm_t, want_rev)
        if want_rev:
            corr, keep = ops.remap_pairs(data['correspondences'].to(dev), rev_s, rev_t)
            data['correspondences'] = corr[:, keep]                            # ordered compaction of the kept columns
        data['src_xyz'], data['tgt_xyz'] = src, tgt
        if src_ov is not None:
            data['src_overlap'] = src_ov
        if tgt_ov is not None:
            data['tgt_overlap'] = tgt_ov
        return data
