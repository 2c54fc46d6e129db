"""Development aid: does one step of the path leave reference cycles behind (tensors that only the cycle collector frees)?
Runs the 64-pair step with the collector disabled, then collects with DEBUG_SAVEALL and prints what was unreachable, and how
the caching allocator's reserved / allocated bytes move over back-to-back steps with the collector off."""
import gc
import os
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import kpconv_config
from kpreg_b200.pipeline import RegistrationPath, result_rows
from bench import make_pairs

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda")
cfg = kpconv_config("3dmatch")
torch.manual_seed(0)
path = RegistrationPath(cfg, index_dtype=torch.int32, weights_threshold=0.85).eval().to(dev)
src, tgt, poses = make_pairs(min(pairs, 8), 1000)
base = len(src)
src = [torch.from_numpy(src[i % base]).to(dev) for i in range(pairs)]
tgt = [torch.from_numpy(tgt[i % base]).to(dev) for i in range(pairs)]
poses = torch.from_numpy(poses[[i % base for i in range(pairs)]]).to(dev)
corr = path(src, tgt, poses)["corr"]
for _ in range(3):
    path(src, tgt, poses, corr=corr)
torch.cuda.synchronize()
gc.collect()
gc.disable()


def stats(tag):
    st = torch.cuda.memory_stats(dev)
    print(f"{tag}: reserved {st['reserved_bytes.all.current'] / 2**30:.3f} GiB, allocated {st['allocated_bytes.all.current'] / 2**30:.3f} GiB, "
          f"cudaMalloc calls {st['num_device_alloc']}", flush=True)


stats("before")
last = None
for i in range(8):
    out = path(src, tgt, poses, corr=corr)
    last = (result_rows(out), out)
    torch.cuda.synchronize()
    stats(f"step {i}")
del last, out
gc.set_debug(gc.DEBUG_SAVEALL)
n = gc.collect()
print("unreachable objects after 8 steps with the collector off:", n)
kinds = Counter(type(o).__name__ for o in gc.garbage)
print(kinds.most_common(15))
tb = sum(o.numel() * o.element_size() for o in gc.garbage if isinstance(o, torch.Tensor))
print(f"tensors among them: {sum(isinstance(o, torch.Tensor) for o in gc.garbage)}, {tb / 2**20:.1f} MiB")
for o in gc.garbage:
    if isinstance(o, torch.Tensor):
        print("  tensor", tuple(o.shape), o.dtype)
gc.set_debug(0)
gc.garbage.clear()
stats("after collect")
