"""kpreg_front_forward on the res2net shapes of the 64-pair 3DMatch workload: us per call and algorithmic GB/s
(x read once + z written once, the x copy included).  Usage: python tools/bench_front.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda")
shapes = [("L0 w28", 2653008, 32, 28), ("L1 w28", 943672, 32, 28), ("L1 w56", 943672, 64, 56), ("L2 w56", 248240, 64, 56)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"{'layer':<8} {'M':>8} {'c_in':>5} {'w':>4} | {'us':>8} {'GB/s':>7}")
for name, m, c, w in shapes:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(m, c, device=dev)
    w1 = (torch.randn(8 * w, c, generator=g) / c ** 0.5).to(dev)
    b1 = (0.1 * torch.randn(8 * w, generator=g)).to(dev)
    wc = (torch.randn(7, w, w, generator=g) / w ** 0.5).to(dev)
    bc = (0.1 * torch.randn(7, w, generator=g)).to(dev)
    pack = ops.FrontPack(w1, b1, wc, bc)
    z = torch.empty(m, 8 * w + c, device=dev)
    for _ in range(3):
        ops.front_forward(x, pack, z, copy_x=True)
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.front_forward(x, pack, z, copy_x=True)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    nb = 4 * m * (c + 8 * w + c)
    print(f"{name:<8} {m:>8} {c:>5} {w:>4} | {t:8.1f} {nb / t / 1e3:7.0f}")
