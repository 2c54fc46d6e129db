"""-m gpu: weighted Kabsch on CUDA vs the reference's compute_rigid_transform (golden) and the oracle.
Tolerance (north star): rotation within 1e-3 deg, translation within 1e-5 m."""
import numpy as np
import pytest
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import synthetic
from kpreg_b200.se3_torch import (compute_rigid_transform, compute_rigid_transform_batch, fast_compute_rigid_transform,
                                  se3_compare)
from gpu_util import cuda

pytestmark = pytest.mark.gpu
ROT_TOL_DEG, TRANS_TOL = 1e-3, 1e-5


def _check(oracle, got, want):
    err = oracle.pose_error(got.detach().cpu(), torch.as_tensor(want))
    assert float(err["rot_deg"].max()) < ROT_TOL_DEG, float(err["rot_deg"].max())
    assert float(err["trans"].max()) < TRANS_TOL, float(err["trans"].max())


def test_matches_reference_golden(oracle, golden_modelnet):
    g = golden_modelnet
    a, b, w = cuda(g["kb_a"]), cuda(g["kb_b"]), cuda(g["kb_w"])
    t = compute_rigid_transform(a, b, w)
    assert t.shape == (6, 3, 4)
    _check(oracle, t, g["kb_T_weighted"])
    _check(oracle, compute_rigid_transform(a, b), g["kb_T_unweighted"])
    _check(oracle, compute_rigid_transform(a, b, torch.zeros_like(w)), g["kb_T_zero"])
    w_fast = w.clone()
    _check(oracle, fast_compute_rigid_transform(a, b, w_fast, 0.85), g["kb_T_fast"])
    # the reference zeroes the caller's weights in place (se3_torch.py:240-242)
    assert torch.equal(w_fast, torch.where(w > 0.85, w, torch.zeros_like(w)))
    # no leading dims
    _check(oracle, compute_rigid_transform(a[0], b[0], w[0]), g["kb_T_weighted"][0])


@pytest.mark.parametrize("seed,n_sets,n_pts,noise", [(0, 6, 1200, 0.01), (1, 1, 3, 0.0), (2, 48, 1400, 0.05), (3, 5, 40000, 0.02)])
def test_matches_oracle(oracle, seed, n_sets, n_pts, noise):
    a, b, w, _ = synthetic.kabsch_inputs(seed, n_sets, n_pts, noise)
    _check(oracle, compute_rigid_transform(cuda(a), cuda(b), cuda(w)), oracle.compute_rigid_transform(a, b, w))


def test_reflection_case_picks_proper_rotation(oracle):
    """Noisy, nearly planar correspondences where V U^T has det < 0: the last column must be flipped."""
    rng = np.random.default_rng(9)
    a = rng.normal(size=(4, 50, 3)).astype(np.float32)
    a[..., 2] *= 1e-3
    b = a.copy()
    b[..., 2] *= -1.0  # a mirror image: best proper rotation is not the reflection
    w = rng.uniform(0.1, 1.0, size=(4, 50)).astype(np.float32)
    got = compute_rigid_transform(cuda(a), cuda(b), cuda(w))
    det = torch.det(got[..., :3, :3].double().cpu())
    assert torch.allclose(det, torch.ones_like(det), atol=1e-5)
    _check(oracle, got, oracle.compute_rigid_transform(a, b, w))


def test_ragged_batch_single_launch(oracle):
    rng = np.random.default_rng(4)
    sets = [synthetic.kabsch_inputs(10 + i, 6, int(n), 0.01) for i, n in enumerate(rng.integers(900, 1500, size=5))]
    got = compute_rigid_transform_batch([cuda(s[0]) for s in sets], [cuda(s[1]) for s in sets], [cuda(s[2]) for s in sets], 0.85)
    assert got.shape == (6, 5, 3, 4)
    for i, s in enumerate(sets):
        _check(oracle, got[:, i], oracle.fast_compute_rigid_transform(s[0], s[1], s[2], 0.85))
    err = se3_compare(got[-1], torch.stack([cuda(s[3]) for s in sets]))
    assert float(err["rot_deg"].max()) < 1.0


def test_asserts_like_reference():
    a = torch.zeros(2, 5, 3, device="cuda")
    with pytest.raises(AssertionError):
        compute_rigid_transform(a, a, torch.full((2, 5), 1.5, device="cuda"))
    with pytest.raises(AssertionError):
        compute_rigid_transform(a, torch.zeros(2, 4, 3, device="cuda"))
