"""Helpers shared by the -m gpu parity tests (CUDA path vs the CPU oracle)."""
import numpy as np
import torch

from test_oracle import LEVEL_KEYS, assert_rows_equal_up_to_ties, check_pyramid, golden_pyramid_3dmatch, _levels  # noqa: F401


def cuda(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def meta_to_numpy(meta):
    return {k: [t.detach().cpu().numpy() for t in v] for k, v in meta.items()}
