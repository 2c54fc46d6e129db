"""Development aid: device time of kpreg_linear_forward for the encoder's layer shapes at the bench workload
(64 stacked 3DMatch-shape pairs), tcgen05 path, with the effective algorithmic bandwidth and fp32-equivalent TF/s.
Environment switches of the library (KPREG_GEMM_STORE, KPREG_GEMM_NO_WIDE3) are read at load time: one process per
setting."""
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


shapes = [(2653008, 32, 224, "", "conv1 L0"), (2653008, 64, 32, "", "unary1 L0"), (2653008, 64, 128, "", "shortcut L0"),
          (2653008, 256, 128, "", "conv3 L0"), (943672, 64, 448, "", "conv1 L1"), (943672, 512, 256, "", "conv3 L1"),
          (943672, 512, 256, "p", "conv3 L1 + shortcut"), (248240, 112, 112, "2", "chain L2"), (248240, 128, 896, "", "conv1 L2"),
          (248240, 1024, 512, "", "conv3 L2"), (60872, 2048, 1024, "", "conv3 L3")]
print(f"KPREG_GEMM_STORE={os.environ.get('KPREG_GEMM_STORE', '1')}")
print(f"{'layer':>22} {'M':>8} {'K':>5} {'N':>5} | {'us':>8} {'GB/s':>7} {'TF/s':>6}")
for m, k, n, flags, name in shapes:
    x = torch.randn(m, k, device="cuda")
    w = torch.randn(n, k, device="cuda")
    shift = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda")
    kw = {}
    extra = 0
    if "2" in flags:
        kw = dict(out2=torch.empty(m, n, device="cuda"), addend=torch.randn(m, n, device="cuda"))
        extra = 2 * m * n
    if "p" in flags:
        kw = dict(post_residual=torch.randn(m, n, device="cuda"), post_act="leaky_relu")
        extra = m * n
    t = timeit(lambda: ops.linear_forward(x, w, None, shift, act="relu", out=out, gemm=1, **kw))
    print(f"{name:>22} {m:>8} {k:>5} {n:>5} | {t:8.1f} {(m * k + m * n + extra) * 4 / t / 1e3:7.0f} {2 * m * k * n / t / 1e6:6.1f}")
    del x, w, out, kw
