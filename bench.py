#!/usr/bin/env python
"""bench.py — pairs/sec through the KPConv registration hot path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one batch of P synthetic 3DMatch-shape pairs per GPU (BASELINE configs[1]: ~20 k points per
cloud after a 2.5 cm voxel grid, 4-level pyramid, neighborhood_limits 40, K = 15) through
subsample pyramid + all neighbour / pool / upsample tables + KPFEncoder forward + weighted Kabsch.

Printed JSON (rank 0):
  value   whole-job pairs/s with the clouds already resident in HBM (device-timed, max over ranks)
  e2e     the same metric through the public API from pinned HOST clouds (H2D of the step's clouds and
          D2H of its poses + errors inside the timed region)
  roofline      dominant kernel family: algorithmic bytes (or flops) / device time (CUDA events recorded
                by the library around its launches, inside the timed region) vs MEASURED_PEAKS.json
  cpu_baseline  (N=1, rank 0) the reference path on the host cores, bounded sample
--impl reference times the reference's own CPU path (compiled reference C++ behind oracle/_ref when
present + the torch-CPU restatement of its encoder / Kabsch) on the same workload, one pair per step.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

# Caching-allocator policy for a steady step time (measured, gpurun_out r3b/r3c): with the default policy the multi-GB
# intermediates of a step are carved out of — and split — whatever large block is free, the free list drifts from step to
# step, and every few dozen steps a 2.7 GB request finds no block: one cudaMalloc in the timed region (3-5 ms of host time
# normally, 30-190 ms in one run of three — the "slow fourth step" of earlier rounds).  Blocks above 128 MB are therefore
# never split: each large size keeps its own blocks and the steady state is reached in the first warm-up steps.
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "max_split_size_mb:128")

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "point-cloud pairs/sec (subsample+neighbours+KPConv+Kabsch)"
UNIT = "pairs/s"
# operand split of the tcgen05 GEMMs (csrc/kpconv_gemm.cu gemm_h2()): fp16 hi + 2^11-scaled fp16 lo by default, 3xTF32 on request
SPLIT_NAME = "3xTF32 split" if os.environ.get("KPREG_GEMM_TF32", "")[:1] == "1" else "fp16 hi/lo split (3 products, fp32 accumulate)"
WORKLOAD = "3DMatch-shape synthetic pairs (~20k pts/cloud, voxel 0.025 m, 4-level KPConv pyramid)"


def make_pairs(n_pairs, seed0):
    from kpreg_b200 import synthetic
    pairs = [synthetic.threedmatch_pair(seed=seed0 + i) for i in range(n_pairs)]
    return [p[0] for p in pairs], [p[1] for p in pairs], np.stack([p[2] for p in pairs])


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md's clocks line), read through NVML (pynvml) by
    the launching thread itself right after the LAST timed step has been enqueued, while the GPU is still executing it.
    Every earlier placement cost one step of the region 50-230 ms: a background `nvidia-smi -lms` or NVML thread stalled CUDA
    calls directly, and even a synchronous query between two steps was followed, three steps later and at that position
    every time, by one slow step (no such step without the queries: the driver appears to do deferred work after an NVML
    clock query).  Fallback without pynvml: one `nvidia-smi` query at the same point."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}  # NVML reason bits

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.nvml, self.handle, self.max_mhz, self.calls = gpu_index, [], None, None, None, 0

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.gpu < len(ids) and ids[self.gpu].isdigit():
                    idx = int(ids[self.gpu])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.sample()  # the first query attaches NVML to the GPU: keep that out of the timed region as well
            self.samples.clear()
        except Exception:
            self.nvml = None

    def sample(self):
        self.calls += 1
        if self.nvml is not None:
            try:
                self.samples.append((float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)),
                                     int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle))))
            except Exception:
                pass
            return
        if self.calls != 1:
            return
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout
            f = [x.strip() for x in out.strip().splitlines()[0].split(",")]
            bits = 0
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    bits |= self.BITS[name]
            self.samples.append((float(f[1]), bits))
            self.max_mhz = float(f[2])
        except Exception:
            pass

    def stop(self):
        got = self.samples
        reasons = sorted(name for name, bit in self.BITS.items() if any(r & bit for _, r in got))
        return {"sm_mhz": float(np.median([c for c, _ in got])) if got else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(got), "source": "nvml" if self.nvml is not None else "nvidia-smi",
                "when": "after the last timed step was enqueued, GPU still executing it"}


def kpconv_work(meta, cfg):
    """Algorithmic bytes / flops of the step's KPConv calls and neighbour queries (SURVEY.md §8d formulas)."""
    n = [int(p.shape[0]) for p in meta["points"]]
    k = cfg.num_kernel_points
    out_dim, in_dim, layer = cfg.first_feats_dim, cfg.in_feats_dim, 0
    gather_bytes = contract_flops = gather_flops = linear_bytes = linear_flops = norm_bytes = 0

    def chain_ok(width, n_layers):
        from kpreg_b200 import kpconv_blocks, ops
        return kpconv_blocks.CHAIN_KERNEL and kpconv_blocks.FUSED_GLUE and ops.chain_supported(width, n_layers)

    def lin(m, k_in, n_out):
        nonlocal linear_bytes, linear_flops
        linear_bytes += 4 * m * (k_in + n_out) + 4 * k_in * n_out
        linear_flops += 2 * m * k_in * n_out

    for name in cfg.architecture:
        strided = "strided" in name
        if name.startswith("simple"):
            c_in, c_out = in_dim, out_dim // 2
        else:
            c_in = c_out = out_dim // 4
        n_s, n_q = n[layer], n[layer + 1] if strided else n[layer]
        h = int(meta["pools"][layer].shape[1] if strided else meta["neighbors"][layer].shape[1])
        # KPConv: idx + points + x + out + weights, each touched once
        gather_bytes += 4 * n_q * h + 12 * (n_s + n_q) + 4 * n_s * c_in + 4 * n_q * c_out + 4 * k * c_in * c_out
        gather_flops += 2 * n_q * k * h * c_in + 12 * n_q * h * k
        contract_flops += 2 * n_q * k * c_in * c_out
        if name.startswith("simple"):
            norm_bytes += 8 * n_q * c_out
        else:
            # unary1, KPConv norm, res2net (conv1, 7 chained, downsample, conv3), shortcut unary (+ norm)
            mid, wid = out_dim // 4, int(out_dim * 14 / 64)
            if in_dim != mid:
                lin(n_s, in_dim, mid)
                norm_bytes += 8 * n_s * mid
            norm_bytes += 8 * n_q * mid
            lin(n_q, mid, 8 * wid)
            if chain_ok(wid, 7):
                # the seven chained layers in one kernel: t read once, the concatenation (+ the copy of the block input) written once
                linear_bytes += 4 * n_q * (2 * 8 * wid + 2 * mid) + 7 * 4 * wid * wid
                linear_flops += 7 * 2 * n_q * wid * wid
            else:
                for _ in range(7):
                    lin(n_q, wid, wid)
            # conv3 and the residual projection as one GEMM over the K-concatenation [cat | x]
            lin(n_q, 8 * wid + mid, out_dim)
            if in_dim != out_dim:
                lin(n_q, in_dim, out_dim)
                norm_bytes += 12 * n_q * out_dim
            else:
                linear_bytes += 4 * n_q * out_dim  # identity shortcut read by conv3's epilogue
        in_dim = out_dim // 2 if name.startswith("simple") else out_dim
        if strided:
            layer += 1
            out_dim *= 2
    query_bytes = 0
    for lvl in range(len(n)):
        w = int(meta["neighbors"][lvl].shape[1])
        query_bytes += 12 * 2 * n[lvl] + 4 * n[lvl] * w
        if lvl + 1 < len(n):
            query_bytes += 12 * (n[lvl] + n[lvl + 1]) + 4 * n[lvl + 1] * int(meta["pools"][lvl].shape[1])
            query_bytes += 12 * (n[lvl] + n[lvl + 1]) + 4 * n[lvl] * int(meta["upsamples"][lvl].shape[1])
    sub_bytes = sum(12 * n[l] + 12 * n[l + 1] for l in range(len(n) - 1))
    return {"kpconv_bytes": gather_bytes, "gather_flops": gather_flops, "contract_flops": contract_flops,
            "query_bytes": query_bytes, "subsample_bytes": sub_bytes, "linear_bytes": linear_bytes,
            "linear_flops": linear_flops, "segment_norm_bytes": norm_bytes}


# ------------------------------------------------------------------------------------------------------
# CPU reference path (oracle): test/bench infrastructure, never the product
# ------------------------------------------------------------------------------------------------------

def _import_reference_encoder():
    """The reference's own KPFEncoder class when /root/reference is present (the build container); None on the GPU box."""
    if not os.path.isdir("/root/reference/models/backbone_kpconv"):
        return None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        cwd = os.getcwd()
        from make_golden import import_reference
        fk, _, _ = import_reference()       # chdirs into the reference tree (its kernel dispositions are CWD-relative)
        os.chdir(cwd)
        return fk
    except Exception as exc:  # noqa: BLE001
        print(f"bench.py: reference import failed ({exc}); using the port", file=sys.stderr)
        return None


class CpuReference:
    """The reference path on the host cores: preprocessing = the UNMODIFIED reference C++ (oracle/_ref) when it was built,
    else the C port; encoder = the reference's own KPFEncoder when /root/reference is importable (build container), else the
    torch-CPU port (oracle/kp_oracle.py); Kabsch = the port of compute_rigid_transform.  `kind` is "reference" only when
    preprocessing AND encoder are the reference's own code."""

    def __init__(self, cfg, state_dict):
        import kp_oracle
        self.o = kp_oracle
        kp_oracle.build()
        self.impl = "ref" if kp_oracle.have_ref() else "port"
        self.cfg = cfg
        self.sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.ref_encoder = None
        fk = _import_reference_encoder()
        if fk is not None:
            cwd = os.getcwd()
            os.chdir("/root/reference")
            try:
                enc = fk.KPFEncoder(cfg, cfg.d_embed)
                enc.load_state_dict(self.sd, strict=True)
                self.ref_encoder = enc.eval()
            finally:
                os.chdir(cwd)
        self.kind = "reference" if (self.impl == "ref" and self.ref_encoder is not None) else "port"

    def describe(self):
        pre = "unmodified reference C++ (oracle/_ref, 1 thread — it is single-threaded by construction)" if self.impl == "ref" else "C port (1 thread)"
        enc = ("the reference's own KPFEncoder (imported from /root/reference)" if self.ref_encoder is not None
               else "torch-CPU port of the reference's KPFEncoder (oracle/kp_oracle.py; /root/reference is absent on this box)")
        return f"preprocess = {pre}; encoder = {enc} on {self.cores} threads; Kabsch = port of compute_rigid_transform"

    def pair(self, src, tgt, pose, seed=0):
        """One pair through preprocess -> encoder -> Kabsch on the host.  Returns stage seconds."""
        from kpreg_b200.pipeline import synthetic_correspondences
        t0 = time.perf_counter()
        meta = self.o.preprocess([src, tgt], self.cfg, impl=self.impl)
        t1 = time.perf_counter()
        x0 = np.ones((meta["points"][0].shape[0], 1), np.float32)
        with torch.no_grad():
            if self.ref_encoder is not None:
                batch = {k: [torch.from_numpy(np.ascontiguousarray(a)) for a in v] for k, v in meta.items()}
                feats, _ = self.ref_encoder(torch.from_numpy(x0), batch)
            else:
                feats, _ = self.o.encoder_forward(self.sd, self.cfg, x0, meta)
        t2 = time.perf_counter()
        lens = [int(v) for v in meta["stack_lengths"][-1]]
        a, b, w = synthetic_correspondences(torch.from_numpy(meta["points"][-1]), lens, torch.from_numpy(pose)[None], seed=seed)
        t3 = time.perf_counter()
        pose_out = self.o.fast_compute_rigid_transform(a[0], b[0], w[0], 0.85)
        t4 = time.perf_counter()
        self.last = {"meta": meta, "feats": feats, "pose": pose_out, "corr": (a[0], b[0], w[0])}
        return {"preprocess": t1 - t0, "encoder": t2 - t1, "kabsch": t4 - t3}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    import kpreg_b200  # noqa: F401
    from kpreg_b200 import kpconv_config
    from kpreg_b200.kpconv import KPFEncoder
    cfg = kpconv_config("3dmatch")
    torch.manual_seed(0)
    np.random.seed(0)
    enc = KPFEncoder(cfg, cfg.d_embed).eval()
    ref = CpuReference(cfg, enc.state_dict())
    src, tgt, poses = make_pairs(1, 1000)
    for _ in range(min(max(args.warmup, 0), 1)):  # bounded: at most one warm-up pair
        ref.pair(src[0], tgt[0], poses[0])
    stages, t_total = [], 0.0
    for _ in range(args.steps):
        st = ref.pair(src[0], tgt[0], poses[0])
        stages.append(st)
        t_total += sum(st.values())
    value = args.steps / t_total
    split = {k: float(np.mean([s[k] for s in stages])) for k in stages[0]}
    sample = (f"{args.steps} steps x 1 pair of the workload; {ref.describe()}; "
              f"stage s/pair {json.dumps({k: round(v, 4) for k, v in split.items()})}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "pairs_per_step": 1},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------

def parity_check(ref, path, out, src_np, tgt_np, poses_np, n_pairs):
    """After the timed region: pair 0 of the measured batch through the CPU reference path (the checker), compared with the
    rows of the GPU batch that belong to that pair — pyramid points bit-exact, encoder features <= 1e-4 (max-norm), pose
    <= 1e-3 deg / 1e-5 m on the same correspondences."""
    import kp_oracle
    ref.pair(src_np[0], tgt_np[0], poses_np[0])
    want = ref.last
    meta = out["meta"]
    points_equal, feat_err = True, None
    for lvl, pts in enumerate(meta["points"]):
        lens = meta["stack_lengths"][lvl].cpu().numpy().astype(np.int64)
        offs = np.concatenate([[0], np.cumsum(lens)])
        rows = np.concatenate([np.arange(offs[0], offs[1]), np.arange(offs[n_pairs], offs[n_pairs + 1])])
        got = pts[torch.from_numpy(rows).to(pts.device)].cpu().numpy()
        points_equal = points_equal and got.shape == want["meta"]["points"][lvl].shape and bool(np.array_equal(got, want["meta"]["points"][lvl]))
        if lvl == len(meta["points"]) - 1:
            f = out["feats"][torch.from_numpy(rows).to(pts.device)].cpu().numpy()
            w = want["feats"].numpy()
            feat_err = float(np.abs(f - w).max() / max(np.abs(w).max(), 1e-30)) if f.shape == w.shape else float("inf")
    # pose: the GPU Kabsch on pair 0's correspondences of the batch vs the port on the same numbers
    a, b, w, offsets = out["corr"]
    lo, hi = int(offsets[0]), int(offsets[6])
    n_c = (hi - lo) // 6
    t_ref = kp_oracle.fast_compute_rigid_transform(a[lo:hi].reshape(6, n_c, 3).cpu(), b[lo:hi].reshape(6, n_c, 3).cpu(),
                                                   w[lo:hi].reshape(6, n_c).cpu(), 0.85)
    perr = kp_oracle.pose_error(out["poses"][:, 0].cpu(), t_ref)
    rot, trans = float(perr["rot_deg"].max()), float(perr["trans"].max())
    ok = bool(points_equal and feat_err is not None and feat_err < 1e-4 and rot < 1e-3 and trans < 1e-5)
    return {"pair": 0, "checker": ref.kind, "points_bit_exact": points_equal, "feature_rel_err": feat_err, "feature_bound": 1e-4,
            "pose_rot_deg_err": rot, "pose_trans_err": trans, "ok": ok}


def other_configs(dev, steps=5):
    """BASELINE configs 1, 3 and 4 (N = 1): single-pair latency of the ModelNet and MCD shapes through pyramid + encoder +
    Kabsch, and the 8-pair training step (forward + backward through the encoder).  Device-timed with CUDA events."""
    import kpreg_b200  # noqa: F401
    from kpreg_b200 import kpconv_config, synthetic
    from kpreg_b200.kpconv import KPFEncoder, Preprocessor
    from kpreg_b200.pipeline import RegistrationPath
    res = {}

    def timed(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    for name, gen in (("modelnet", synthetic.modelnet_pair), ("mcd", synthetic.mcd_pair)):
        cfg = kpconv_config(name)
        torch.manual_seed(0)
        np.random.seed(0)
        path = RegistrationPath(cfg, index_dtype=torch.int32, weights_threshold=0.85).eval().to(dev)
        src, tgt, pose = gen(seed=7)
        s, t, p = [torch.from_numpy(src).to(dev)], [torch.from_numpy(tgt).to(dev)], torch.from_numpy(pose)[None].to(dev)
        out = path(s, t, p)
        ms = timed(lambda: path(s, t, p), steps)
        res[f"{name}_single_pair"] = {"ms_per_pair": ms, "pairs_per_s": 1000.0 / ms,
                                      "points_per_level": [int(x.shape[0]) for x in out["meta"]["points"]],
                                      "neighbor_widths": [int(x.shape[1]) for x in out["meta"]["neighbors"]]}
        del path, out
    cfg = kpconv_config("3dmatch")
    torch.manual_seed(0)
    np.random.seed(0)
    pairs = [synthetic.threedmatch_pair(seed=100 + i) for i in range(8)]
    pts = [torch.from_numpy(p[0]).to(dev) for p in pairs] + [torch.from_numpy(p[1]).to(dev) for p in pairs]
    pre = Preprocessor(cfg, index_dtype=torch.int32)
    enc = KPFEncoder(cfg, cfg.d_embed).train().to(dev)
    meta = pre(pts)
    x0 = torch.ones((meta["points"][0].shape[0], 1), device=dev)

    def train_step():
        enc.zero_grad(set_to_none=True)
        y, _ = enc(x0, meta)
        y.square().mean().backward()

    ms = timed(train_step, 3)
    res["train_step_8_pairs"] = {"ms_per_step": ms, "pairs_per_s": 8000.0 / ms,
                                 "what": "forward + backward through all 11 KPConv ops and 3 max_pools of the encoder (BatchNorm in training mode), pyramid precomputed"}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=64, help="pairs per GPU per (sub-)batch")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --pairs per GPU per step; strong: --global-pairs per step sharded round-robin over the GPUs (SURVEY §8d config 5)")
    ap.add_argument("--global-pairs", type=int, default=512)
    ap.add_argument("--gemm", type=int, default=None, help="contractions: 0 fp32 CUDA cores, 1 tcgen05 split-operand GEMM (default; fp16 hi/lo split, or 3xTF32 under KPREG_GEMM_TF32=1)")
    ap.add_argument("--no-fused-glue", action="store_true", help="run the block glue on stock PyTorch ops")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the ModelNet / MCD / training-step side measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # NVML is attached (and queried once) FIRST, seconds before anything is timed: the driver does some deferred work ~2.5 s
    # after a process attaches — with the sampler started in front of the ramp-up that was one 60-230 ms step at a fixed
    # position (the fourth) of the timed region in every second run.  The samples themselves are taken at the end of the region.
    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("KPREG_BENCH_NO_CLOCKS") else None
    if sampler:
        sampler.start()

    import torch.distributed as dist
    import kpreg_b200  # noqa: F401
    from kpreg_b200 import _lib, kpconv_blocks, kpconv_config
    from kpreg_b200.pipeline import RegistrationPath, gather_results, result_rows, shard_pairs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    json_out = sys.stdout
    if world > 1:
        # rank 0 prints exactly one JSON line on stdout.  NCCL writes its own log lines (version banner, NCCL_DEBUG=INFO
        # topology / rank lines) to file descriptor 1 from C: point fd 1 at stderr for the life of the process — nothing is
        # silenced, the lines can still be read there — and keep the original stdout for the JSON line alone.
        sys.stdout.flush()
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    if args.gemm is not None:
        kpconv_blocks.DEFAULT_GEMM = args.gemm
    if args.no_fused_glue:
        kpconv_blocks.FUSED_GLUE = False

    cfg = kpconv_config("3dmatch")
    torch.manual_seed(0)
    np.random.seed(0)
    path = RegistrationPath(cfg, index_dtype=torch.int32, weights_threshold=0.85).eval().to(dev)

    # the step's pairs of this rank, as sub-batches of at most --pairs pairs
    if args.scaling == "weak":
        n_global = args.pairs * world                      # every rank gets its own P pairs (global pair = rank + world * i)
        src_np, tgt_np, poses_np = make_pairs(args.pairs, 1000 + 100 * rank)
        my_batches = [list(range(args.pairs))]
    else:
        n_global = args.global_pairs                       # fixed global set, pair g -> rank g mod world
        n_unique = min(n_global, 64)                       # (64 distinct synthetic pairs, cycled: generation is host time)
        src_np, tgt_np, poses_np = make_pairs(n_unique, 1000)
        mine = [g % n_unique for g in shard_pairs(n_global, rank, world)]
        my_batches = [mine[i:i + args.pairs] for i in range(0, len(mine), args.pairs)]
    n_local = sum(len(b) for b in my_batches)

    class Batch:
        """One sub-batch: clouds resident on the device, and the same clouds in ONE pinned host buffer (a single
        host-to-device copy per sub-batch in the end-to-end measurement)."""

        def __init__(self, ids):
            self.n = len(ids)
            clouds = [src_np[i] for i in ids] + [tgt_np[i] for i in ids]
            self.lens = [int(c.shape[0]) for c in clouds]
            self.host = torch.from_numpy(np.concatenate(clouds, 0)).pin_memory()
            self.poses_host = torch.from_numpy(np.stack([poses_np[i] for i in ids])).pin_memory()
            dev_all = self.host.to(dev)
            parts = list(torch.split(dev_all, self.lens))
            self.src_dev, self.tgt_dev = parts[:self.n], parts[self.n:]
            self.poses_dev = self.poses_host.to(dev)
            self.h2d_bytes = self.host.numel() * 4 + self.poses_host.numel() * 4
            # decoder-shaped correspondences: in the model they are the decoder's output; the synthetic stand-ins are built
            # once, outside the timed region, from this batch's own coarse level
            self.corr = path(self.src_dev, self.tgt_dev, self.poses_dev)["corr"]

    batches = [Batch(ids) for ids in my_batches]
    h2d_bytes = sum(b.h2d_bytes for b in batches)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def finish(rows_list, last_out):
        rows = torch.cat(rows_list, 0) if len(rows_list) > 1 else rows_list[0]
        if args.scaling == "strong" and world > 1:
            per = (n_global + world - 1) // world
            padded = torch.zeros((per, rows.shape[1]), dtype=rows.dtype, device=rows.device)
            padded[:rows.shape[0]] = rows
            table = torch.empty((world * per, rows.shape[1]), dtype=rows.dtype, device=rows.device)
            dist.all_gather_into_tensor(table, padded)
            return table, last_out
        return gather_results(rows, n_local * world, rank, world), last_out

    def step_resident():
        rows, out = [], None
        for b in batches:
            out = path(b.src_dev, b.tgt_dev, b.poses_dev, corr=b.corr)
            rows.append(result_rows(out))
        return finish(rows, out)

    host_table = [None]

    def step_e2e():
        rows, out = [], None
        for b in batches:
            parts = list(torch.split(b.host.to(dev, non_blocking=True), b.lens))   # ONE copy of the sub-batch's clouds
            out = path(parts[:b.n], parts[b.n:], b.poses_host.to(dev, non_blocking=True), corr=b.corr)
            rows.append(result_rows(out))
        table, out = finish(rows, out)
        # D2H of every pair's pose + errors into pinned host memory, stream-ordered (inside the step's CUDA events); the host
        # does not block on it, so a step's launches are not serialised behind the previous step's read-back
        if host_table[0] is None or host_table[0].shape != table.shape:
            host_table[0] = torch.empty(table.shape, dtype=table.dtype).pin_memory()
        host_table[0].copy_(table, non_blocking=True)
        return host_table[0], out

    def timed(fn, steps, warmup, profile, between=None):
        if profile:
            _lib.profile(True)
        last = None
        for _ in range(warmup):
            last = fn()  # (kept while the next step runs, exactly as in the timed loop: the caching allocator sees one pattern)
            flush.fill_(1)
        torch.cuda.synchronize()
        # The per-family CUDA events are recorded in the first `prof_steps` steps of the timed region only: thousands of
        # pending event pairs slow the stream down (measured: 7 000 scopes over 20 steps cost 5 ms per step, 1 750 nothing).
        prof_steps = min(5, steps) if profile else 0
        if profile:
            per_step = sum(v[1] for v in _lib.profile_read().values()) // max(warmup, 1) + 8
            prof_steps = max(1, min(prof_steps, 1800 // per_step))  # (a strong-scaling step is several sub-batches: fewer steps)
            if os.environ.get("KPREG_BENCH_PROF_STEPS"):
                prof_steps = max(1, min(steps, int(os.environ["KPREG_BENCH_PROF_STEPS"])))
            _lib.profile_reserve(per_step * (prof_steps + 1))  # their events exist before the timed region starts
        # Long-lived Python objects (modules, cached packs, the batches) leave the collector's working set: a full
        # collection in the middle of a step otherwise stalls the launching thread for 0.1-0.2 s (seen as one 70-240 ms
        # step in ten, always at the same call count) while the GPU drains; the collector is paused for the timed steps.
        gc.collect()
        gc.freeze()
        gc.disable()  # (reference counting still frees every tensor of a step at once; only cycle detection waits for the region's end)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches0 = _lib.launch_count()
        if profile:
            _lib.profile(True)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        host_ms = []
        mem0 = torch.cuda.memory_stats(dev)
        for i in range(steps):
            if profile and i == prof_steps:
                _lib.profile(False)  # stops recording; the records stay readable
            starts[i].record()
            t_h = time.perf_counter()
            last = fn()
            host_ms.append(round((time.perf_counter() - t_h) * 1e3, 2))
            ends[i].record()
            flush.fill_(i & 1)  # evict L2 between timed steps (outside the events)
            if between is not None and i == steps - 1:
                # clock samples while the GPU is still busy with the last timed step's tail (about half a step is enqueued
                # but not executed at this point) — see ClockSampler for why not earlier
                for _ in range(5):
                    between()
                    time.sleep(0.003)
        torch.cuda.synchronize()
        gc.enable()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        fam = _lib.profile_read() if profile else None
        if profile:
            _lib.profile(False)
            fam = {k: (v[0] * steps / prof_steps, v[1] * steps / prof_steps) for k, v in fam.items()}  # scaled to `steps` steps
        per_step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
        timed.last_steps = [round(v, 2) for v in per_step_ms]
        # diagnostics of the region (a slow step with a long host time is a host stall; cudaMalloc / cudaFree inside it is the allocator)
        mem1 = torch.cuda.memory_stats(dev)
        timed.last_host = host_ms
        timed.last_alloc = {k: int(mem1.get(k, 0) - mem0.get(k, 0)) for k in ("num_device_alloc", "num_device_free", "num_alloc_retries",
                                                                              "reserved_bytes.all.current")}
        ms = sum(per_step_ms)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), _lib.launch_count() - launches0, fam, last

    # clock / allocator / page-cache ramp-up of a fresh box: run the step untimed for ~2 s before the W warm-up steps
    t_ramp = time.perf_counter()
    ramp_s = 0.0 if os.environ.get("KPREG_BENCH_NO_RAMP") else 2.0  # (profilers count launches: no time-based loop)
    keep = out_r = None
    while time.perf_counter() - t_ramp < ramp_s:
        # local work only: the iteration count is time-based and differs per rank, so NO collective in here.  The previous
        # step's result stays alive while the next one runs, as in the timed loop: measured, the caching allocator otherwise
        # meets a new live-set pattern in the timed region and called cudaMalloc in its fourth step in every run (3-5 ms of host
        # time, and 30-170 ms in one run of five: the "slow fourth step").
        for b in batches:
            out_r = path(b.src_dev, b.tgt_dev, b.poses_dev, corr=b.corr)
            keep = (result_rows(out_r), out_r)
        flush.fill_(1)
        torch.cuda.synchronize()
    del keep, out_r

    ms_total, launches, fam, last = timed(step_resident, args.steps, args.warmup, profile=True,
                                          between=sampler.sample if sampler else None)
    steps_resident, host_resident, alloc_resident = timed.last_steps, timed.last_host, timed.last_alloc
    clocks = sampler.stop() if sampler else None
    keep = None
    for _ in range(0 if os.environ.get("KPREG_BENCH_NO_RAMP") else 4):  # the e2e step's own allocation pattern (H2D staging), untimed
        keep = step_e2e()
        flush.fill_(1)
    del keep
    ms_e2e, _, _, last_e2e = timed(step_e2e, args.steps, args.warmup, profile=False)
    steps_e2e, host_e2e, alloc_e2e = timed.last_steps, timed.last_host, timed.last_alloc

    value = n_global * args.steps / (ms_total / 1000.0) if args.scaling == "strong" else n_local * world * args.steps / (ms_total / 1000.0)
    e2e_value = value * ms_total / ms_e2e
    table, out = last
    d2h_bytes = int(last_e2e[0].numel() * 4)

    if rank == 0:
        pk = peaks()
        scale_work = n_local / batches[-1].n       # the families' times cover every sub-batch; the work formulas the last one
        work = {k: v * scale_work for k, v in kpconv_work(out["meta"], cfg).items()}
        fam_ms = {k: v[0] / args.steps for k, v in fam.items()}
        fam_n = {k: v[1] / args.steps for k, v in fam.items()}
        top = max(("kpconv_gather", "kpconv_contract", "grid_query", "subsample", "linear"), key=lambda k: fam_ms[k])
        nbytes = {"kpconv_gather": work["kpconv_bytes"], "grid_query": work["query_bytes"], "subsample": work["subsample_bytes"],
                  "linear": work["linear_bytes"], "kpconv_contract": None}[top]
        names = {"kpconv_gather": "k_kpconv_gather_mma + k_kpconv_c1 (KPConv gather + influence + aggregation)",
                 "grid_query": "k_grid_query_tq / k_grid_query (radius neighbours)", "subsample": "subsample_batch (all kernels)",
                 "linear": f"k_gemm_tc + k_res2net_front (block Linear layers: tcgen05 {SPLIT_NAME} GEMMs; conv1 + chain of the narrow res2net units in one tcgen05 kernel)",
                 "kpconv_contract": f"k_gemm_tc (KPConv contraction [Nq,K*Cin]x[K*Cin,Cout], tcgen05 {SPLIT_NAME})"}
        if top == "kpconv_contract":
            ach = work["contract_flops"] / (fam_ms[top] * 1e-3) / 1e12
            roof = {"kernel": names[top], "bound": "tensor", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": ach / pk["tflops"], "traffic": None}
        else:
            ach = nbytes / (fam_ms[top] * 1e-3) / 1e9
            roof = {"kernel": names[top], "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / pk["hbm_gbs"], "traffic": None}
        # every family against its own bound, for the record (TF32 peak: profiles/r2_tf32_peak.json, cuBLAS 8192^3 on this pool)
        tf32 = None
        tpk = os.path.join(ROOT, "profiles", "r2_tf32_peak.json")
        if os.path.exists(tpk):
            tf32 = json.load(open(tpk)).get("tf32_tflops_sustained")
        hbm = pk["hbm_gbs"]
        fams = {
            "kpconv_gather_GBs": work["kpconv_bytes"] / (fam_ms["kpconv_gather"] * 1e-3) / 1e9,
            "kpconv_gather_fp32_TFLOPs": work["gather_flops"] / (fam_ms["kpconv_gather"] * 1e-3) / 1e12,
            "kpconv_contract_TFLOPs": work["contract_flops"] / max(fam_ms["kpconv_contract"], 1e-9) / 1e9,
            "linear_GBs": work["linear_bytes"] / max(fam_ms["linear"], 1e-9) / 1e6,
            "linear_TFLOPs": work["linear_flops"] / max(fam_ms["linear"], 1e-9) / 1e9,
            "grid_query_GBs": work["query_bytes"] / max(fam_ms["grid_query"], 1e-9) / 1e6,
            "segment_norm_GBs": work["segment_norm_bytes"] / max(fam_ms["segment_norm"], 1e-9) / 1e6,
            "subsample_GBs": work["subsample_bytes"] / max(fam_ms["subsample"], 1e-9) / 1e6,
        }
        fams["frac_of_hbm"] = {k[:-4]: round(fams[k] / hbm, 4) for k in ("kpconv_gather_GBs", "linear_GBs", "grid_query_GBs", "segment_norm_GBs", "subsample_GBs")}
        if tf32:
            # split operands: three tensor-core products per fp32-equivalent product (the fp16 split issues kind::f16 MMAs, whose
            # dense peak is twice TF32's — the TF32 figure is kept as the common denominator)
            fams["kpconv_contract_frac_of_tf32_peak"] = {"fp32_equivalent": round(fams["kpconv_contract_TFLOPs"] / tf32, 4),
                                                          "issued_3x": round(3 * fams["kpconv_contract_TFLOPs"] / tf32, 4), "tf32_peak_TFLOPs": tf32}
        roof["families"] = fams
        # DRAM traffic of the family from the committed ncu capture of this build (dram__bytes_read.sum + dram__bytes_write.sum
        # over one step's launches), scaled to this run's pairs per step
        tpath = next((q for q in (os.path.join(ROOT, "profiles", f) for f in ("r3_dram_traffic.json", "r2_dram_traffic.json")) if os.path.exists(q)),
                     os.path.join(ROOT, "profiles", "r3_dram_traffic.json"))
        fam_kernels = {"linear": ("k_gemm_tc", "k_chain", "k_res2net_front"), "kpconv_contract": ("k_gemm_tc",),
                       "kpconv_gather": ("k_kpconv_gather_mma", "k_kpconv_c1"), "grid_query": ("k_grid_query",)}.get(top)
        if os.path.exists(tpath) and fam_kernels:
            tr = json.load(open(tpath))
            fs = [tr["families"][k] for k in fam_kernels if k in tr["families"]]
            if fs:
                roof["traffic"] = sum(f["dram_read_MB"] + f["dram_write_MB"] for f in fs) * 1e6 * n_local / tr["pairs"]
                roof["traffic_note"] = (f"bytes per step, all {' + '.join(fam_kernels)} launches (ncu capture at {tr['pairs']} pairs/step, "
                                        f"profiles/{os.path.basename(tpath)}, scaled to {n_local}; k_gemm_tc also serves the KPConv contraction); "
                                        "achieved / algorithmic figures are per step as well")
        if roof["traffic"] is None:
            roof["traffic_note"] = "profiles/r3_dram_traffic.json is missing"
        roof["peak_source"] = pk["src"] + " (MEASURED_PEAKS.json)" if pk["src"] == "measured" else "fallback"
        roof["per_step_ms"] = {k: round(v, 4) for k, v in fam_ms.items()}
        roof["launches_per_step"] = {k: v for k, v in fam_n.items()}
        roof["share_of_step"] = round(fam_ms[top] / (ms_total / args.steps), 4)
        roof["untimed_residue_ms"] = round(ms_total / args.steps - sum(fam_ms.values()), 4)
        roof["algorithmic_per_step"] = work

        cpu, parity = None, None
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(cfg, path.kpf_encoder.state_dict())
            ids = my_batches[-1]
            parity = parity_check(ref, path, out, [src_np[i] for i in ids], [tgt_np[i] for i in ids], [poses_np[i] for i in ids], len(ids))
            n_s, t_s, stages = 2, 0.0, []
            for i in range(n_s):
                j = ids[i % len(ids)]
                st = ref.pair(src_np[j], tgt_np[j], poses_np[j])
                stages.append(st)
                t_s += sum(st.values())
            split = {k: round(float(np.mean([s[k] for s in stages])), 4) for k in stages[0]}
            cpu = {"value": n_s / t_s, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                   "sample": f"{n_s} pairs of the same workload after 1 warm-up pair (the parity check's); {ref.describe()}; "
                             f"stage s/pair {json.dumps(split)}"}
        extra = None
        if world == 1 and not args.no_other_configs:
            del batches
            torch.cuda.empty_cache()
            extra = other_configs(dev)

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": n_local, "global_pairs_per_step": n_global if args.scaling == "strong" else n_local * world,
                       "sub_batch_pairs": args.pairs,
                       "points_per_level": [int(p.shape[0]) for p in out["meta"]["points"]],
                       "neighbor_widths": [int(t.shape[1]) for t in out["meta"]["neighbors"]],
                       "kpconv_contraction": f"tcgen05 {SPLIT_NAME}" if kpconv_blocks.DEFAULT_GEMM == 1 else "fp32-cuda-core",
                       "block_glue": "fused CUDA (tcgen05 linear + segment norm)" if kpconv_blocks.FUSED_GLUE else "PyTorch ops",
                       "parallelism": f"pairs sharded over {world} GPU(s), no data-path collective; all-gather of [P,14] poses+errors per step",
                       "kabsch_inputs": "decoder-shaped synthetic correspondences (the decoder is out of scope), built once from the batch's "
                                        "own coarse level outside the timed region and resident on the device",
                       "l2": "256 MiB write between timed steps (outside the CUDA events)",
                       "allocator": "PYTORCH_CUDA_ALLOC_CONF=" + os.environ.get("PYTORCH_CUDA_ALLOC_CONF", ""),
                       "ramp_up": "2 s of untimed steps before the W warm-up steps (fresh-box clocks / allocator), 4 untimed e2e steps before the e2e warm-up"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / args.steps, "h2d_copies_per_step": 2 * len(my_batches)},
            "gpu_launches": int(launches),
            "step_ms": {"resident": steps_resident, "e2e": steps_e2e, "host_enqueue_resident": host_resident, "host_enqueue_e2e": host_e2e,
                        "allocator_in_region": {"resident": alloc_resident, "e2e": alloc_e2e}},
            "roofline": roof,
            "cpu_baseline": cpu,
            "parity_check": parity,
            "other_configs": extra,
            "final_layer_pose_error": {"rot_deg_max": float(table[:, 12].max()), "trans_max": float(table[:, 13].max()),
                                       "note": "against the generating pose through 1 cm synthetic correspondence noise — not a parity figure"},
        }
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
