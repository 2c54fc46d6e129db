"""Drop-in replacements for the reference's two CPython extension modules.

    from kpreg_b200.cpp_wrappers import cpp_subsampling, cpp_neighbors
    s_points, s_len = cpp_subsampling.subsample_batch(points, batches, sampleDl=0.05, max_p=0, verbose=0)
    neighbors      = cpp_neighbors.batch_query(queries, supports, q_batches, s_batches, radius=0.0625)

mirror ``grid_subsampling.subsample_batch`` (reference cpp_wrappers/cpp_subsampling/wrapper.cpp:62-333,
format ``"OO|$OOfsii"``) and ``radius_neighbors.batch_query`` (cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238,
format ``"OOOO|$f"``): two / four positional arrays, the rest keyword-only, float32 points [N,3],
int32 batch lengths, RuntimeError on malformed input or an empty result.

Inputs may be numpy arrays / CPU tensors (as in the reference; outputs are then numpy arrays, and the
host<->device copies happen here) or CUDA tensors (outputs stay on the device).  The computation
itself always runs on the GPU — there is no CPU path.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from . import ops


def _to_device(a, dtype, what: str):
    """Returns (cuda tensor, was_cuda)."""
    if isinstance(a, torch.Tensor):
        if a.is_cuda:
            return a.to(dtype).contiguous(), True
        return a.detach().to(dtype).contiguous().cuda(non_blocking=True), False
    try:
        arr = np.ascontiguousarray(a, dtype=np.float32 if dtype == torch.float32 else np.int32)
    except Exception as exc:  # same failure mode as PyArray_FROM_OTF returning NULL
        raise RuntimeError(f"Error converting {what} to numpy arrays of type "
                           f"{'float32' if dtype == torch.float32 else 'int32'}") from exc
    return torch.from_numpy(arr).cuda(non_blocking=True), False


def _check_points(t: torch.Tensor, what: str) -> None:
    if t.dim() != 2 or t.shape[1] != 3:
        raise RuntimeError(f"Wrong dimensions : {what}.shape is not (N, 3)")


def _check_batches(t: torch.Tensor, what: str) -> None:
    if t.dim() != 1:
        raise RuntimeError(f"Wrong dimensions : {what}.shape is not (B,)")


def subsample_batch(points, batches, *, features=None, classes=None, sampleDl=0.1, method="barycenters",
                    max_p=0, verbose=0):
    """(s_points f32 [M,3], s_len i32 [B]) — voxel-grid barycentres per cloud, reference order."""
    if features is not None or classes is not None:
        # the repository only ever calls the points-only branch (finegrained_kpconv.py:371)
        raise NotImplementedError("subsample_batch: features / classes are not used by the registration path")
    pts, on_dev = _to_device(points, torch.float32, "points")
    lens, _ = _to_device(batches, torch.int32, "batches")
    _check_points(pts, "points")
    _check_batches(lens, "batches")
    out, counts = ops.subsample(pts, lens, float(sampleDl), int(max_p))
    host = counts.cpu()  # the result shape is data dependent: one small read-back
    if int(host[-1]) != 0:
        raise RuntimeError("subsample_batch: voxel grid too large to index (sampleDl too small for the extent)")
    m = int(host[-2])
    if m < 1:
        raise RuntimeError("Error")  # wrapper.cpp:266-270
    s_points, s_len = out[:m], counts[:-2]
    if on_dev:
        return s_points.clone(), s_len.clone()
    return s_points.cpu().numpy(), host[:-2].numpy().copy()


def batch_query(queries, supports, q_batches, s_batches, *, radius=0.1):
    """int32 [Nq, max_count]: per query the same-cloud supports with d2 < radius^2, ascending, padded
    with the total support count."""
    q, on_dev = _to_device(queries, torch.float32, "query points")
    s, _ = _to_device(supports, torch.float32, "support points")
    qb, _ = _to_device(q_batches, torch.int32, "query batches")
    sb, _ = _to_device(s_batches, torch.int32, "support batches")
    _check_points(q, "queries")
    _check_points(s, "supports")
    _check_batches(qb, "q_batches")
    _check_batches(sb, "s_batches")
    if qb.shape[0] != sb.shape[0]:
        raise RuntimeError("Wrong number of batch elements: different for queries and supports")
    grid = ops.CellGrid(s, sb, float(radius))
    # pass 1: widest row (nothing is written at width 0); pass 2: fill exactly that width
    _, _, stats = grid.query(q, qb, float(radius), 0)
    host = stats.cpu()
    if int(host[1]) != 0:
        raise RuntimeError("batch_query: cell grid too large to index")
    width = int(host[0])
    if q.shape[0] < 1 or width < 1:
        raise RuntimeError("Error")  # wrapper.cpp:201-205
    idx, _, _ = grid.query(q, qb, float(radius), width)
    return idx if on_dev else idx.cpu().numpy()


# module-shaped handles, so `cpp_subsampling.subsample_batch(...)` / `cpp_neighbors.batch_query(...)` read
# exactly like the reference's imports (finegrained_kpconv.py:12-15)
cpp_subsampling = SimpleNamespace(subsample_batch=subsample_batch)
cpp_neighbors = SimpleNamespace(batch_query=batch_query)
grid_subsampling = cpp_subsampling
radius_neighbors = cpp_neighbors
