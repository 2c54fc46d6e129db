"""Development aid: one benchmark step under torch.profiler — where does the step's GPU time go
(our kernels vs the PyTorch glue of the encoder blocks) and how much of the step is host-bound."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

import kpreg_b200  # noqa: F401
from kpreg_b200 import kpconv_blocks, kpconv_config
from kpreg_b200.pipeline import RegistrationPath, result_rows
from bench import make_pairs

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
if len(sys.argv) > 2:
    kpconv_blocks.DEFAULT_GEMM = int(sys.argv[2])
cfg = kpconv_config("3dmatch")
torch.manual_seed(0)
np.random.seed(0)
dev = torch.device("cuda")
path = RegistrationPath(cfg, index_dtype=torch.int32, weights_threshold=0.85).eval().to(dev)
src, tgt, poses = make_pairs(pairs, 1000)
src = [torch.from_numpy(a).to(dev) for a in src]
tgt = [torch.from_numpy(a).to(dev) for a in tgt]
poses = torch.from_numpy(poses).to(dev)
for _ in range(3):
    result_rows(path(src, tgt, poses))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    result_rows(path(src, tgt, poses))
torch.cuda.synchronize()
print(f"wall ms/step: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ({pairs} pairs)")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    result_rows(path(src, tgt, poses))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
