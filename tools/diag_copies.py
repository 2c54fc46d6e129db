"""Development aid: which PyTorch ops (copies, casts, element-wise kernels) are left in one step of the path, grouped by input
shape with the Python stack that issued them — how the strided copies of the wide res2net units were found."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from torch.profiler import ProfilerActivity, profile
import kpreg_b200
from kpreg_b200 import kpconv_config
from kpreg_b200.pipeline import RegistrationPath, result_rows
from bench import make_pairs
pairs = 16
cfg = kpconv_config("3dmatch")
dev = torch.device("cuda")
path = RegistrationPath(cfg, index_dtype=torch.int32, weights_threshold=0.85).eval().to(dev)
src, tgt, poses = make_pairs(pairs, 1000)
src = [torch.from_numpy(a).to(dev) for a in src]; tgt = [torch.from_numpy(a).to(dev) for a in tgt]
poses = torch.from_numpy(poses).to(dev)
corr = path(src, tgt, poses)["corr"]
for _ in range(2): path(src, tgt, poses, corr=corr)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    path(src, tgt, poses, corr=corr)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6) if e.key in ("aten::copy_", "aten::contiguous", "aten::clone", "aten::to", "aten::_to_copy", "aten::cat", "aten::fill_", "aten::add", "aten::mul", "aten::index", "aten::zeros", "aten::ones")]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:14]:
    print(e.key, e.count, round(e.device_time_total), e.input_shapes)
    for s in e.stack[:6]: print("     ", s)
