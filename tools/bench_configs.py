"""Development aid: single-pair latency of the hot path (pyramid / encoder) for the three BASELINE shapes, with the
per-family device time of the preprocessing kernels."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import _lib, kpconv_config, synthetic
from kpreg_b200.kpconv import KPFEncoder, Preprocessor

for name, gen, kw in (("modelnet", synthetic.modelnet_pair, {}), ("3dmatch", synthetic.threedmatch_pair, {}),
                      ("mcd", synthetic.mcd_pair, {})):
    cfg = kpconv_config(name)
    torch.manual_seed(0)
    np.random.seed(0)
    src, tgt, _ = gen(seed=7, **kw)
    pts = [torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()]
    pre = Preprocessor(cfg, index_dtype=torch.int32)
    enc = KPFEncoder(cfg, cfg.d_embed).eval().cuda()
    for _ in range(3):
        meta = pre(pts)
        with torch.no_grad():
            enc(torch.ones((meta["points"][0].shape[0], 1), device="cuda"), meta)
    torch.cuda.synchronize()
    _lib.profile(True)
    t0 = time.perf_counter()
    n_it = 5
    for _ in range(n_it):
        meta = pre(pts)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(n_it):
        with torch.no_grad():
            enc(torch.ones((meta["points"][0].shape[0], 1), device="cuda"), meta)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    fam = _lib.profile_read()
    _lib.profile(False)
    print(f"{name}: levels {[int(p.shape[0]) for p in meta['points']]} widths {[int(t.shape[1]) for t in meta['neighbors']]}")
    print(f"   preprocess {1e3 * (t1 - t0) / n_it:.2f} ms/pair, encoder {1e3 * (t2 - t1) / n_it:.2f} ms/pair; device ms/pair: "
          + ", ".join(f"{k} {v[0] / n_it:.3f}" for k, v in fam.items() if v[1]))

# BASELINE config 4: training step (forward + backward through KPConv / max_pool, autograd-mode glue) on 8 stacked
# 3DMatch-shape pairs per GPU
cfg = kpconv_config("3dmatch")
torch.manual_seed(0)
np.random.seed(0)
pairs = [synthetic.threedmatch_pair(seed=100 + i) for i in range(8)]
pts = [torch.from_numpy(p[0]).cuda() for p in pairs] + [torch.from_numpy(p[1]).cuda() for p in pairs]
pre = Preprocessor(cfg, index_dtype=torch.int32)
enc = KPFEncoder(cfg, cfg.d_embed).train().cuda()
meta = pre(pts)
x0 = torch.ones((meta["points"][0].shape[0], 1), device="cuda")


def train_step():
    enc.zero_grad(set_to_none=True)
    out, _ = enc(x0, meta)
    out.square().mean().backward()


for _ in range(2):
    train_step()
torch.cuda.synchronize()
_lib.profile(True)
t0 = time.perf_counter()
for _ in range(3):
    train_step()
torch.cuda.synchronize()
t1 = time.perf_counter()
fam = _lib.profile_read()
_lib.profile(False)
print(f"training step, 8 pairs (levels {[int(p.shape[0]) for p in meta['points']]}): {1e3 * (t1 - t0) / 3:.1f} ms fwd+bwd through the encoder; "
      "device ms/step: " + ", ".join(f"{k} {v[0] / 3:.2f}" for k, v in fam.items() if v[1]))
