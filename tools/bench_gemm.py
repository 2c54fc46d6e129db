"""Development aid: device time of kpreg_linear_forward for the encoder's layer shapes (8 stacked 3DMatch-shape
pairs), tcgen05 path vs fp32 CUDA cores vs torch.mm, with the effective A+C bandwidth."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import ops


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3  # us


shapes = [(331626, 480, 32, "kpconv L0"), (331626, 64, 32, "unary1 L0"), (331626, 32, 224, "res2net conv1 L0"),
          (331626, 28, 28, "res2net chain L0"), (331626, 224, 128, "res2net conv3 L0"), (331626, 64, 128, "shortcut L0"),
          (117959, 960, 64, "kpconv L1"), (117959, 448, 256, "conv3 L1"), (117959, 56, 56, "chain L1"),
          (31030, 1920, 128, "kpconv L2"), (31030, 896, 512, "conv3 L2"), (7609, 3840, 256, "kpconv L3"),
          (7609, 1792, 1024, "conv3 L3")]
print(f"{'shape':>28} {'M':>7} {'K':>5} {'N':>5} | {'tc us':>8} {'GB/s':>7} {'TF/s':>6} | {'simt us':>8} | {'torch us':>8}")
for m, k, n, name in shapes:
    x = torch.randn(m, k, device="cuda")
    w = torch.randn(n, k, device="cuda")
    out = torch.empty(m, n, device="cuda")
    t_tc = timeit(lambda: ops.linear_forward(x, w, out=out, gemm=1))
    t_simt = timeit(lambda: ops.linear_forward(x, w, out=out, gemm=0), n=3) if m * k * n < 3e10 else float("nan")
    t_torch = timeit(lambda: torch.mm(x, w.t(), out=out))
    gbs = (m * k + m * n) * 4 / t_tc / 1e3
    tfs = 2 * m * k * n / t_tc / 1e6
    print(f"{name:>28} {m:>7} {k:>5} {n:>5} | {t_tc:8.1f} {gbs:7.0f} {tfs:6.1f} | {t_simt:8.1f} | {t_torch:8.1f}")
