"""cuBLAS TF32 dense peak on this GPU ("TF32 peak must be measured"), the denominator for the
3xTF32 contractions: torch.matmul fp32 8192^3 with TF32 tensor cores allowed, best of 10 (burst) and back to back for
4 s (sustained) — the same procedure MEASURED_PEAKS.json documents for bf16.  Also bf16 for a same-run cross-check.
Usage: python tools/measure_tf32_peak.py > profiles/rNN_tf32_peak.json"""
import json
import time

import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
flops = 2.0 * n ** 3
out = {"gpu": torch.cuda.get_device_name(0), "n": n, "how": "torch.matmul 8192^3, CUDA events; burst = best of 10, sustained = back to back for 4 s"}
for name, dtype in (("tf32", torch.float32), ("bf16", torch.bfloat16)):
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        a @ b
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, reps = time.perf_counter(), 0
    s.record()
    while time.perf_counter() - t0 < 4.0:
        for _ in range(20):
            a @ b
        reps += 20
        torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    out[f"{name}_tflops"] = round(flops / (best * 1e-3) / 1e12, 1)
    out[f"{name}_tflops_sustained"] = round(flops * reps / (s.elapsed_time(e) * 1e-3) / 1e12, 1)
print(json.dumps(out, indent=1))
