"""Deterministic encoder weights shared by the fixture generator (tests/golden/make_golden_r2.py, which runs the
REFERENCE's KPFEncoder) and the parity tests (which load the same values into this repo's KPFEncoder): every tensor
of the state dict is a pure function of its NAME and shape, so the fixture only has to store the kernel points (they
come from the reference's disposition files + a random rotation) and the reference's output features.

BatchNorm1d layers get non-trivial running statistics and affine parameters so that the folded Linear+BN path of the
res2net units is exercised with numbers that matter."""
import zlib

import numpy as np
import torch


def seeded_state_dict(state_dict, salt: int = 0):
    """{name: tensor} with the same names / shapes / dtypes as ``state_dict``; ``kernel_points`` entries are kept."""
    out = {}
    for name, ref in state_dict.items():
        if name.endswith("kernel_points") or not ref.is_floating_point():
            out[name] = ref.detach().clone()
            continue
        rng = np.random.default_rng(zlib.crc32(name.encode()) + 1000003 * salt)
        shape = tuple(ref.shape)
        if name.endswith("running_mean"):
            v = 0.2 * rng.standard_normal(shape)
        elif name.endswith("running_var"):
            v = 0.5 + rng.random(shape)
        elif ".bn" in name and name.endswith(".weight") or name.endswith("downsample.1.weight"):
            v = 0.8 + 0.4 * rng.random(shape)
        elif name.endswith(".bias"):
            v = 0.1 * rng.standard_normal(shape)
        elif name.endswith("KPConv.weights"):            # [K, c_in, c_out]
            v = rng.standard_normal(shape) * (1.5 / np.sqrt(shape[0] * shape[1]))
        else:                                             # nn.Linear weight [out, in]
            v = rng.standard_normal(shape) * (1.0 / np.sqrt(shape[-1]))
        out[name] = torch.from_numpy(np.asarray(v, np.float32)).to(ref.dtype)
    return out
