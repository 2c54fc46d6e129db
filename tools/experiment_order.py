"""Experiment: does a 3-D blocked (tile-major) processing order of the query rows speed up the gather-type kernels
(KPConv aggregate, max_pool) over the cell-sorted row-major order the pyramid hands out?  The order only permutes the
processing sequence, never the results."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import kpreg_b200  # noqa: F401
from kpreg_b200 import _lib, kpconv_config
from kpreg_b200.kpconv import KPFEncoder, Preprocessor
from bench import make_pairs

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = kpconv_config("3dmatch")
torch.manual_seed(0)
np.random.seed(0)
dev = torch.device("cuda")
src, tgt, _ = make_pairs(pairs, 1000)
pts = [torch.from_numpy(a).to(dev) for a in src] + [torch.from_numpy(a).to(dev) for a in tgt]
meta = Preprocessor(cfg, index_dtype=torch.int32)(pts)
enc = KPFEncoder(cfg, cfg.d_embed).eval().to(dev)
x0 = torch.ones((meta["points"][0].shape[0], 1), device=dev)


def blocked_order(points, lens, cell, tile):
    lens = lens.long()
    cloud = torch.repeat_interleave(torch.arange(lens.shape[0], device=dev), lens)
    c = torch.floor((points - points.min(0).values) / cell).long()
    t = c // tile
    r = c % tile
    key = cloud
    for v, n in ((t[:, 2], 4096), (t[:, 1], 4096), (t[:, 0], 4096), (r[:, 2], tile), (r[:, 1], tile), (r[:, 0], tile)):
        key = key * n + v
    return torch.argsort(key, stable=True).to(torch.int32)


def run(label):
    for _ in range(2):
        with torch.no_grad():
            enc(x0, meta)
    torch.cuda.synchronize()
    _lib.profile(True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        with torch.no_grad():
            enc(x0, meta)
    e.record()
    torch.cuda.synchronize()
    fam = _lib.profile_read()
    _lib.profile(False)
    print(f"{label}: encoder {s.elapsed_time(e) / 3:.2f} ms; gather {fam['kpconv_gather'][0] / 3:.2f}, max_pool {fam['max_pool'][0] / 3:.2f}, "
          f"contract {fam['kpconv_contract'][0] / 3:.2f}, linear {fam['linear'][0] / 3:.2f}, segnorm {fam['segment_norm'][0] / 3:.2f}")


run("cell-sorted row-major order")
orig = meta["orders"]
r = cfg.first_subsampling_dl * cfg.conv_radius
for tile in (2, 4, 8):
    meta["orders"] = [blocked_order(meta["points"][l], meta["stack_lengths"][l], r * 2 ** l, tile) for l in range(len(meta["points"]))]
    run(f"blocked order, tiles of {tile}^3 cells")
meta["orders"] = [None] * len(meta["points"])
run("no order (input order)")
meta["orders"] = orig
run("cell-sorted row-major order (again)")
