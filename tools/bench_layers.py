"""Development aid: per-call device time of every operator launch of one encoder pass at the bench workload
(default 64 stacked 3DMatch-shape pairs): Linear layers (shape, epilogue flags), segment norms, KPConv, max-pool,
with the algorithmic bandwidth of each call.  CUDA events around each call, median of --reps passes."""
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import kpconv_config, ops
from kpreg_b200.pipeline import RegistrationPath
from bench import make_pairs

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
only = sys.argv[3] if len(sys.argv) > 3 else ""

records = []  # (name, desc, bytes, start, end)
active = [False]


def wrap(name, fn, describe):
    def inner(*a, **kw):
        if not active[0]:
            return fn(*a, **kw)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = fn(*a, **kw)
        e.record()
        desc, nbytes = describe(out, *a, **kw)
        records.append((name, desc, nbytes, s, e))
        return out
    return inner


def d_linear(_res, x, weight, col_scale=None, col_shift=None, residual=None, act=None, slope=0.1, out=None, out2=None,
             addend=None, gemm=1, **kw):
    m, k = x.shape
    n = weight.shape[0]
    flags = ("R" if residual is not None else "") + ("2" if out2 is not None or kw.get("out2") is not None else "") + (act or "")[:1]
    nb = 4 * m * (k + n) + (4 * m * n if residual is not None else 0) + (8 * m * n if out2 is not None else 0)
    return f"M={m} K={k} N={n} {flags}", nb


def d_norm(_res, x, lens, residual=None, **kw):
    n, c = x.shape
    return f"N={n} C={c}{' R' if residual is not None else ''}", 4 * n * c * (3 if residual is None else 4)


def d_kpconv(_res, q_pts, s_pts, idx, x, weights, *a, **kw):
    k, ci, co = weights.shape
    nq, h = idx.shape
    return f"Nq={nq} H={h} Cin={ci} Cout={co}", 4 * nq * h + 12 * (nq + s_pts.shape[0]) + 4 * x.shape[0] * ci + 4 * nq * co


def d_pool(_res, x, idx, *a, **kw):
    nq, h = idx.shape
    return f"Nq={nq} H={h} C={x.shape[1]}", 4 * nq * h + 4 * x.shape[0] * x.shape[1] + 4 * nq * x.shape[1]


def d_chain(_res, t, pack, z, x_copy=None):
    m = t.shape[0]
    g = (pack.n_layers + 1) * pack.width
    return f"M={m} w={pack.width} L={pack.n_layers}", 4 * m * (2 * g + (2 * x_copy.shape[1] if x_copy is not None else 0))


def d_front(_res, x, pack, z, copy_x=False):
    m, c = x.shape
    return f"M={m} c_in={c} w={pack.width} G={pack.n_groups}", 4 * m * (c + pack.n_groups * pack.width + (c if copy_x else 0))


def d_pair(_res, x1, x2, weight_cat, *a, **kw):
    m, k1 = x1.shape
    k2, n = x2.shape[1], weight_cat.shape[0]
    return f"M={m} K={k1}+{k2} N={n} pair", 4 * m * (k1 + k2 + n)


ops.chain_forward = wrap("chain", ops.chain_forward, d_chain)
ops.linear_pair_forward = wrap("linear", ops.linear_pair_forward, d_pair)
ops.front_forward = wrap("front", ops.front_forward, d_front)
ops.linear_forward = wrap("linear", ops.linear_forward, d_linear)
ops.segment_norm = wrap("segnorm", ops.segment_norm, d_norm)
ops.kpconv_forward = wrap("kpconv", ops.kpconv_forward, d_kpconv)
ops.max_pool_forward = wrap("max_pool", ops.max_pool_forward, d_pool)

cfg = kpconv_config("3dmatch")
torch.manual_seed(0)
np.random.seed(0)
dev = torch.device("cuda")
path = RegistrationPath(cfg, index_dtype=torch.int32, weights_threshold=0.85).eval().to(dev)
base = min(pairs, 8)
src, tgt, poses = make_pairs(base, 1000)
src = [torch.from_numpy(src[i % base]).to(dev) for i in range(pairs)]
tgt = [torch.from_numpy(tgt[i % base]).to(dev) for i in range(pairs)]
poses = torch.from_numpy(poses[[i % base for i in range(pairs)]]).to(dev)
for _ in range(2):
    path(src, tgt, poses)
torch.cuda.synchronize()
table = OrderedDict()
for r in range(reps):
    records.clear()
    active[0] = True
    path(src, tgt, poses)
    active[0] = False
    torch.cuda.synchronize()
    for i, (name, desc, nb, s, e) in enumerate(records):
        table.setdefault((i, name, desc, nb), []).append(s.elapsed_time(e) * 1e3)
tot = {}
print(f"{'#':>3} {'op':>8} {'shape':<42} {'us':>9} {'GB/s':>7}")
for (i, name, desc, nb), ts in table.items():
    t = float(np.median(ts))
    tot[name] = tot.get(name, 0.0) + t
    if only and only != name:
        continue
    print(f"{i:>3} {name:>8} {desc:<42} {t:9.1f} {nb / t / 1e3:7.0f}")
print("totals (ms):", {k: round(v / 1e3, 2) for k, v in tot.items()})
