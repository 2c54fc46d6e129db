"""Generates tests/golden/*.npz by running the REFERENCE's own code (test infrastructure).

Run in the build container (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed is the unmodified reference:
  * C++ core (grid_subsampling.cpp / neighbors.cpp / nanoflann) through oracle/_ref/libkpref.so,
    injected where finegrained_kpconv.py expects its CPython modules (cpp_subsampling / cpp_neighbors,
    finegrained_kpconv.py:12-15 — the shipped CPython glue does not build against numpy >= 2);
  * Python ``Preprocessor`` (CPU), ``KPFEncoder``, ``KPConv``, ``max_pool`` and
    ``compute_rigid_transform`` imported from /root/reference with MinkowskiEngine / pytorch3d / the
    ``models`` package __init__ stubbed (they are imported at module level but unused on this path)
    and an attr-dict in place of easydict — the recipe of SURVEY.md Appendix A.
Inputs are the seeded synthetic clouds of kpreg_b200.synthetic.  The fixtures pin the oracle
(tests/test_oracle.py) and the CUDA path (tests/test_gpu_*.py).
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import kp_oracle  # noqa: E402
import kpreg_b200  # noqa: E402,F401
from kpreg_b200 import kpconv_config, synthetic  # noqa: E402


def import_reference():
    for name in ["MinkowskiEngine", "pytorch3d", "pytorch3d.ops"]:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorch3d.ops"].packed_to_padded = None
    sys.modules["pytorch3d.ops"].ball_query = None
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = pkg
    sys.path[:0] = [REF, os.path.join(REF, "models")]
    os.chdir(REF)  # load_kernels looks for kernels/dispositions relative to the CWD
    import backbone_kpconv.finegrained_kpconv as fk
    import backbone_kpconv.finegrained_kpconv_blocks as fb
    from utils.se3_torch import compute_rigid_transform

    class _Sub:
        @staticmethod
        def subsample_batch(points, batches, sampleDl=0.1, max_p=0, verbose=0):
            return kp_oracle.subsample_batch(points, batches, sampleDl=sampleDl, max_p=max_p, impl="ref")

    class _Nb:
        @staticmethod
        def batch_query(queries, supports, q_batches, s_batches, radius=0.1):
            return kp_oracle.batch_query(queries, supports, q_batches, s_batches, radius=radius, impl="ref")

    fk.cpp_subsampling = _Sub
    fk.cpp_neighbors = _Nb
    return fk, fb, compute_rigid_transform


def main():
    kp_oracle.build()
    fk, fb, ref_rigid = import_reference()
    out = {}

    # ---------------- ModelNet-shape pair: pyramid + encoder (reduced width so the fixture stays small)
    cfg = kpconv_config("modelnet", first_feats_dim=64)
    src, tgt, pose = synthetic.modelnet_pair(seed=1)
    pts = [torch.from_numpy(src), torch.from_numpy(tgt)]
    meta = fk.Preprocessor(cfg)(pts)
    out["mn_src"], out["mn_tgt"], out["mn_pose"] = src, tgt, pose
    for key in ("points", "neighbors", "pools", "upsamples", "stack_lengths"):
        for lvl, t in enumerate(meta[key]):
            out[f"mn_{key}_{lvl}"] = t.numpy()

    torch.manual_seed(7)
    np.random.seed(7)
    enc = fk.KPFEncoder(cfg, 256)
    # give the BatchNorm1d layers of the res2net units non-trivial statistics
    gen = torch.Generator().manual_seed(11)
    for m in enc.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(0.2 * torch.randn(m.num_features, generator=gen))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=gen))
            m.weight.data.copy_(0.8 + 0.4 * torch.rand(m.num_features, generator=gen))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=gen))
    enc.eval()
    sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    feats0 = torch.ones(meta["points"][0].shape[0], 1)
    with torch.no_grad():
        y, skips = enc(feats0, meta)
    out["mn_enc_out"] = y.numpy()
    for i, s in enumerate(skips):
        out[f"mn_enc_skip_{i}"] = s.numpy()
    for k, v in sd.items():
        out["mn_sd::" + k] = v.numpy()

    # gradients of sum(encoder output) w.r.t. the first two KPConv weight tensors (training-mode BN)
    enc.train()
    enc.zero_grad()
    y_tr, _ = enc(feats0, meta)
    y_tr.sum().backward()
    out["mn_enc_out_train"] = y_tr.detach().numpy()
    out["mn_grad_kp0"] = enc.encoder_blocks[0].KPConv.weights.grad.numpy()
    out["mn_grad_kp1"] = enc.encoder_blocks[1].KPConv.weights.grad.numpy()
    enc.eval()

    # ---------------- single KPConv + max_pool calls (all influence / aggregation modes)
    torch.manual_seed(3)
    np.random.seed(3)
    q_pts, s_pts = meta["points"][1], meta["points"][0]
    idx = meta["pools"][0]
    x = torch.randn(s_pts.shape[0], 24)
    out["op_x"] = x.numpy()
    for infl in ("linear", "gaussian", "constant"):
        for agg in ("sum", "closest"):
            conv = fb.KPConv(15, 3, 24, 40, 0.12, 0.165, KP_influence=infl, aggregation_mode=agg)
            with torch.no_grad():
                o = conv(q_pts, s_pts, idx, x)
            out[f"op_{infl}_{agg}_w"] = conv.weights.detach().numpy()
            out[f"op_{infl}_{agg}_kp"] = conv.kernel_points.detach().numpy()
            out[f"op_{infl}_{agg}_out"] = o.numpy()
    out["op_maxpool"] = fb.max_pool(x, idx).numpy()
    # autograd of one KPConv (linear/sum) for the backward kernels
    conv = fb.KPConv(15, 3, 24, 40, 0.12, 0.165)
    xg = x.clone().requires_grad_(True)
    g = torch.randn(q_pts.shape[0], 40)
    conv(q_pts, s_pts, idx, xg).backward(g)
    out["bw_w"], out["bw_kp"], out["bw_g"] = conv.weights.detach().numpy(), conv.kernel_points.detach().numpy(), g.numpy()
    out["bw_dx"], out["bw_dw"] = xg.grad.numpy(), conv.weights.grad.numpy()
    xg2 = x.clone().requires_grad_(True)
    fb.max_pool(xg2, idx).backward(g[:, :24].contiguous())
    out["bw_maxpool_dx"] = xg2.grad.numpy()

    # ---------------- weighted Kabsch
    a, b, w, kpose = synthetic.kabsch_inputs(seed=5, n_sets=6, n_pts=600)
    out["kb_a"], out["kb_b"], out["kb_w"], out["kb_pose"] = a, b, w, kpose
    out["kb_T_weighted"] = ref_rigid(torch.from_numpy(a), torch.from_numpy(b), torch.from_numpy(w)).numpy()
    out["kb_T_unweighted"] = ref_rigid(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    w_thr = np.where(w > 0.85, w, 0.0).astype(np.float32)  # fast_compute_rigid_transform :240-242 (CUDA-only there)
    out["kb_T_fast"] = ref_rigid(torch.from_numpy(a), torch.from_numpy(b), torch.from_numpy(w_thr)).numpy()
    out["kb_T_zero"] = ref_rigid(torch.from_numpy(a), torch.from_numpy(b), torch.zeros(6, 600)).numpy()

    np.savez_compressed(os.path.join(HERE, "modelnet_pair.npz"), **out)

    # ---------------- 3DMatch-shape pair: pyramid tables only (compressed ints), and subsample outputs
    cfg3 = kpconv_config("3dmatch")
    src, tgt, pose = synthetic.threedmatch_pair(seed=2, n_raw=9000)  # ~6k pts/cloud keeps the file small
    meta3 = fk.Preprocessor(cfg3)([torch.from_numpy(src), torch.from_numpy(tgt)])
    out3 = {"src": src, "tgt": tgt, "pose": pose}
    for key in ("points", "neighbors", "pools", "upsamples", "stack_lengths"):
        for lvl, t in enumerate(meta3[key]):
            arr = t.numpy()
            if key in ("neighbors", "pools", "upsamples"):
                arr = arr.astype(np.int32)
            if key == "points" and lvl == 0:
                continue  # = cat(src, tgt)
            out3[f"{key}_{lvl}"] = arr
    np.savez_compressed(os.path.join(HERE, "threedmatch_small_pyramid.npz"), **out3)

    # ---------------- parameter names / shapes of the full-size encoders (checkpoint compatibility)
    names = {}
    for name in ("3dmatch", "modelnet"):
        c = kpconv_config(name)
        e = fk.KPFEncoder(c, c.d_embed)
        names[name] = {k: list(v.shape) for k, v in e.state_dict().items()}
    with open(os.path.join(HERE, "encoder_state_dict_shapes.json"), "w") as fh:
        json.dump(names, fh, indent=0, sort_keys=True)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
