"""One benchmark step inside a cudaProfilerStart/Stop window, for `ncu --profile-from-start off ...`:
3 untimed steps (allocator, split-weight and chain-pack caches warm), then exactly one step of P stacked 3DMatch-shape
pairs through pyramid + encoder + Kabsch.  Usage: ncu ... python tools/ncu_step.py [pairs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import kpconv_config
from kpreg_b200.pipeline import RegistrationPath, result_rows
from bench import make_pairs

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
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
torch.cuda.profiler.start()
result_rows(path(src, tgt, poses))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"profiled one step of {pairs} pairs")
