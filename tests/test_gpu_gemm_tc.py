"""-m gpu: the tcgen05 (3xTF32) contraction vs the fp32 CUDA-core contraction and the CPU oracle.
The split-operand scheme must hold the same 1e-4 relative tolerance as the fp32 path."""
import numpy as np
import pytest
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import ops
from conftest import rel_err
from gpu_util import cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c_in,c_out", [(32, 32), (64, 64), (128, 128), (256, 256), (24, 40), (32, 28), (12, 200)])
def test_tc_contraction_matches_fp32_and_oracle(oracle, golden_modelnet, c_in, c_out):
    g = golden_modelnet
    rng = np.random.default_rng(c_in * 7 + c_out)
    q, s, idx = g["mn_points_0"], g["mn_points_0"], g["mn_neighbors_0"]
    x = rng.normal(size=(s.shape[0], c_in)).astype(np.float32)
    w = (rng.normal(size=(15, c_in, c_out)) / np.sqrt(15 * c_in)).astype(np.float32)
    kp = g["op_linear_sum_kp"] * 0.5
    args = (cuda(q), cuda(s), cuda(idx), cuda(x), cuda(w), cuda(kp), 0.06)
    simt = ops.kpconv_forward(*args, gemm=0)
    tc = ops.kpconv_forward(*args, gemm=1)
    torch.cuda.synchronize()
    assert rel_err(tc.cpu().numpy(), simt.cpu().numpy()) < 2e-6  # 3xTF32 is fp32-grade, far inside 1e-4
    want = oracle.kpconv_forward(q, s, idx, x, w, kp, 0.06, dtype=torch.float64)
    assert rel_err(tc.cpu().numpy(), want.numpy()) < 1e-5
    assert rel_err(simt.cpu().numpy(), want.numpy()) < 1e-5


def test_tc_contraction_many_row_tiles(oracle):
    """M spanning many 128-row tiles with a ragged tail, wide K (several pipeline rounds)."""
    rng = np.random.default_rng(0)
    n = 128 * 37 + 19
    pts = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    idx = rng.integers(0, n + 1, size=(n, 9)).astype(np.int32)  # includes shadow indices (= n)
    x = rng.normal(size=(n, 64)).astype(np.float32)
    w = rng.normal(size=(15, 64, 64)).astype(np.float32)
    kp = rng.uniform(-0.3, 0.3, size=(15, 3)).astype(np.float32)
    args = (cuda(pts), cuda(pts), cuda(idx), cuda(x), cuda(w), cuda(kp), 1.5)
    simt = ops.kpconv_forward(*args, gemm=0)
    tc = ops.kpconv_forward(*args, gemm=1)
    torch.cuda.synchronize()
    assert rel_err(tc.cpu().numpy(), simt.cpu().numpy()) < 2e-6
