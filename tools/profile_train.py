"""Development aid: the 8-pair training step (BASELINE config 4: forward + backward through the encoder) under
torch.profiler — which kernels carry the step.  Usage: python tools/profile_train.py [pairs]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

import kpreg_b200  # noqa: F401
from kpreg_b200 import kpconv_config, synthetic
from kpreg_b200.kpconv import KPFEncoder, Preprocessor

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = kpconv_config("3dmatch")
torch.manual_seed(0)
np.random.seed(0)
pairs = [synthetic.threedmatch_pair(seed=100 + i) for i in range(n_pairs)]
pts = [torch.from_numpy(p[0]).cuda() for p in pairs] + [torch.from_numpy(p[1]).cuda() for p in pairs]
pre = Preprocessor(cfg, index_dtype=torch.int32)
enc = KPFEncoder(cfg, cfg.d_embed).train().cuda()
meta = pre(pts)
x0 = torch.ones((meta["points"][0].shape[0], 1), device="cuda")


def train_step():
    enc.zero_grad(set_to_none=True)
    out, _ = enc(x0, meta)
    out.square().mean().backward()


for _ in range(3):
    train_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    train_step()
torch.cuda.synchronize()
print(f"training step, {n_pairs} pairs: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms wall")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    train_step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=80))
