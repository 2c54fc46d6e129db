"""TEST INFRASTRUCTURE — CPU oracle, not product code.

CPU restatement of the reference's KPConv hot path.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module; the
product package never does (tests/test_capi.py::test_product_package_never_touches_the_oracle enforces it).

Two native back ends sit behind the same functions:

* ``impl="port"`` — ``oracle/kp_oracle.c``: this repo's plain-C restatement.
* ``impl="ref"``  — ``oracle/_ref/libkpref.so``: the UNMODIFIED reference C++ core compiled from
  ``/root/reference`` behind ``oracle/ref_shim.cpp`` (built here, travels to the GPU box as a
  prebuilt ``.so``; the reference tree itself does not).

Parity pinning: the reference has no tests or golden vectors.  The port is
pinned against ``impl="ref"`` and against ``tests/golden/*.npz``, which were produced by importing
the reference's own Python modules (``tests/golden/make_golden.py``).

Reference citations are relative to ``/root/reference``.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: Dict[str, ctypes.CDLL] = {}

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)


def build(quiet: bool = True) -> None:
    """Compile kp_oracle.c and (when /root/reference is present) oracle/_ref/libkpref.so."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libkpref.so"))


def _lib(impl: str) -> ctypes.CDLL:
    if impl in _LIBS:
        return _LIBS[impl]
    if impl == "port":
        path = os.path.join(_HERE, "libkporacle.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.kporacle_subsample_batch.restype = ctypes.c_int
        lib.kporacle_subsample_batch.argtypes = [_f32p, ctypes.c_int, _i32p, ctypes.c_int,
                                                 ctypes.c_float, ctypes.c_int, _f32p, _i32p]
        lib.kporacle_batch_query.restype = ctypes.c_int
        lib.kporacle_batch_query.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int, _i32p, _i32p,
                                             ctypes.c_int, ctypes.c_float,
                                             ctypes.POINTER(_i32p), ctypes.c_char_p]
        lib.kporacle_free.argtypes = [ctypes.c_void_p]
    elif impl == "ref":
        path = os.path.join(_HERE, "_ref", "libkpref.so")
        if not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/libkpref.so is not built (needs /root/reference)")
        lib = ctypes.CDLL(path)
        lib.kpref_subsample_batch.restype = ctypes.c_int
        lib.kpref_subsample_batch.argtypes = [_f32p, ctypes.c_int, _i32p, ctypes.c_int,
                                              ctypes.c_float, ctypes.c_int,
                                              ctypes.POINTER(_f32p), _i32p]
        lib.kpref_batch_query.restype = ctypes.c_int
        lib.kpref_batch_query.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int, _i32p, _i32p,
                                          ctypes.c_int, ctypes.c_float, ctypes.POINTER(_i32p)]
        lib.kpref_free.argtypes = [ctypes.c_void_p]
    else:
        raise ValueError(impl)
    _LIBS[impl] = lib
    return lib


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.int32)


# --------------------------------------------------------------------------------------------
# Native ops (cpp_wrappers)
# --------------------------------------------------------------------------------------------

def subsample_batch(points, batches, sampleDl: float = 0.1, max_p: int = 0, impl: str = "port"):
    """cpp_subsampling.subsample_batch (cpp_wrappers/cpp_subsampling/wrapper.cpp:62-333).

    Returns (s_points f32 [M,3], s_len i32 [B]).
    """
    p, b = _f32(points), _i32(batches)
    n, nb = p.shape[0], b.shape[0]
    lens = np.zeros(nb, np.int32)
    lib = _lib(impl)
    if impl == "port":
        out = np.zeros((max(n, 1), 3), np.float32)
        m = lib.kporacle_subsample_batch(p.ctypes.data_as(_f32p), n, b.ctypes.data_as(_i32p), nb,
                                         ctypes.c_float(sampleDl), int(max_p),
                                         out.ctypes.data_as(_f32p), lens.ctypes.data_as(_i32p))
        if m < 0:
            raise RuntimeError("oracle subsample failed")
        return out[:m].copy(), lens
    ptr = _f32p()
    m = lib.kpref_subsample_batch(p.ctypes.data_as(_f32p), n, b.ctypes.data_as(_i32p), nb,
                                  ctypes.c_float(sampleDl), int(max_p), ctypes.byref(ptr),
                                  lens.ctypes.data_as(_i32p))
    out = np.ctypeslib.as_array(ptr, shape=(max(m, 1), 3))[:m].copy()
    lib.kpref_free(ptr)
    return out, lens


def batch_query(queries, supports, q_batches, s_batches, radius: float = 0.1, impl: str = "port",
                return_ties: bool = False):
    """cpp_neighbors.batch_query (cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238).

    Returns i32 [Nq, max_count]; rows are ascending-d2, padded with Ns_total.
    ``return_ties`` (port only) additionally returns a bool [Nq] marking rows holding two equal
    d2 values, where the reference's unstable sort leaves the order unspecified.
    """
    q, s, qb, sb = _f32(queries), _f32(supports), _i32(q_batches), _i32(s_batches)
    nq, ns, nb = q.shape[0], s.shape[0], qb.shape[0]
    lib = _lib(impl)
    ptr = _i32p()
    ties = None
    if impl == "port":
        ties = np.zeros(max(nq, 1), np.uint8)
        w = lib.kporacle_batch_query(q.ctypes.data_as(_f32p), nq, s.ctypes.data_as(_f32p), ns,
                                     qb.ctypes.data_as(_i32p), sb.ctypes.data_as(_i32p), nb,
                                     ctypes.c_float(radius), ctypes.byref(ptr),
                                     ties.ctypes.data_as(ctypes.c_char_p))
        free = lib.kporacle_free
    else:
        w = lib.kpref_batch_query(q.ctypes.data_as(_f32p), nq, s.ctypes.data_as(_f32p), ns,
                                  qb.ctypes.data_as(_i32p), sb.ctypes.data_as(_i32p), nb,
                                  ctypes.c_float(radius), ctypes.byref(ptr))
        free = lib.kpref_free
    out = np.ctypeslib.as_array(ptr, shape=(max(nq * w, 1),))[:nq * w].copy().reshape(nq, w)
    free(ptr)
    if return_ties:
        if ties is None:
            raise ValueError("return_ties needs impl='port'")
        return out, ties[:nq].astype(bool)
    return out


def neighbors_kpconv(queries, supports, q_batches, s_batches, radius, max_neighbors, impl="port"):
    """batch_neighbors_kpconv (models/backbone_kpconv/finegrained_kpconv.py:248-263)."""
    nb = batch_query(queries, supports, q_batches, s_batches, radius=radius, impl=impl)
    return nb[:, :max_neighbors] if max_neighbors > 0 else nb


# --------------------------------------------------------------------------------------------
# Pyramid schedule
# --------------------------------------------------------------------------------------------

def preprocess(pts: Sequence, cfg, impl: str = "port") -> Dict[str, List[np.ndarray]]:
    """Preprocessor.forward (models/backbone_kpconv/finegrained_kpconv.py:303-419), numpy out.

    points f32, neighbors/pools/upsamples int64, stack_lengths int32 (int64 placeholder at the
    last level, as in the reference :390-393).
    """
    limits = cfg.neighborhood_limits
    r_normal = cfg.first_subsampling_dl * cfg.conv_radius
    arch = list(cfg.architecture)
    lens = np.array([int(p.shape[0]) for p in pts], np.int32)
    cur = np.concatenate([_f32(p) for p in pts], 0)
    out = {k: [] for k in ("points", "neighbors", "pools", "upsamples", "stack_lengths")}
    layer, layer_blocks = 0, []
    for i, block in enumerate(arch):
        if "global" in block or "upsample" in block:
            break
        strided = "pool" in block or "strided" in block
        if not strided:
            layer_blocks.append(block)
            if i < len(arch) - 1 and "upsample" not in arch[i + 1]:
                continue
        if layer_blocks:
            r = r_normal
            if any("deformable" in b for b in layer_blocks[:-1]):
                r = r_normal * cfg.deform_radius / cfg.conv_radius
            conv_i = neighbors_kpconv(cur, cur, lens, lens, r, limits[layer], impl).astype(np.int64)
        else:
            conv_i = np.zeros((0, 1), np.int64)
        if strided:
            dl = 2 * r_normal / cfg.conv_radius
            pool_p, pool_b = subsample_batch(cur, lens, sampleDl=dl, impl=impl)
            r = r_normal * cfg.deform_radius / cfg.conv_radius if "deformable" in block else r_normal
            pool_i = neighbors_kpconv(pool_p, cur, pool_b, lens, r, limits[layer], impl).astype(np.int64)
            up_i = neighbors_kpconv(cur, pool_p, lens, pool_b, 2 * r, limits[layer], impl).astype(np.int64)
        else:
            pool_i = np.zeros((0, 1), np.int64)
            pool_p = np.zeros((0, 3), np.float32)
            pool_b = np.zeros((0,), np.int64)
            up_i = np.zeros((0, 1), np.int64)
        out["points"].append(cur)
        out["neighbors"].append(conv_i)
        out["pools"].append(pool_i)
        out["upsamples"].append(up_i)
        out["stack_lengths"].append(lens)
        cur, lens = pool_p, pool_b
        r_normal *= 2
        layer += 1
        layer_blocks = []
    return out


# --------------------------------------------------------------------------------------------
# KPConv operator and encoder blocks (torch CPU, fp32 like the reference; fp64 on request)
# --------------------------------------------------------------------------------------------

def kpconv_forward(q_pts, s_pts, neighb_inds, x, weights, kernel_points, KP_extent: float,
                   KP_influence: str = "linear", aggregation_mode: str = "sum",
                   dtype=torch.float32) -> torch.Tensor:
    """KPConv.forward, rigid branch (models/backbone_kpconv/finegrained_kpconv_blocks.py:265-401).

    Shadow support point at 1e6 (:296), influence (:345-362), 'closest' one-hot (:365-369), zero
    shadow feature row (:375), [n,K,H]x[n,H,Cin] then per-kernel-point [Cin,Cout] contraction
    (:381-393), and the fork's normalisation by the number of neighbours whose feature SUM is
    positive, clamped to >= 1 (:396-399).
    """
    q = torch.as_tensor(q_pts).to(dtype).cpu()
    s = torch.as_tensor(s_pts).to(dtype).cpu()
    idx = torch.as_tensor(neighb_inds).long().cpu()
    xf = torch.as_tensor(x).to(dtype).cpu()
    w = torch.as_tensor(weights).to(dtype).cpu()
    kp = torch.as_tensor(kernel_points).to(dtype).cpu()
    s = torch.cat([s, torch.full((1, 3), 1e6, dtype=dtype)], 0)
    rel = s[idx] - q[:, None, :]                                  # [n, H, 3]
    d2 = ((rel[:, :, None, :] - kp[None, None]) ** 2).sum(-1)     # [n, H, K]
    if KP_influence == "constant":
        infl = torch.ones_like(d2)
    elif KP_influence == "linear":
        infl = torch.clamp(1 - torch.sqrt(d2) / KP_extent, min=0.0)
    elif KP_influence == "gaussian":
        sigma = KP_extent * 0.3
        infl = torch.exp(-d2 / (2 * sigma ** 2 + 1e-9))
    else:
        raise ValueError("Unknown influence function type (config.KP_influence)")
    infl = infl.transpose(1, 2)                                   # [n, K, H]
    if aggregation_mode == "closest":
        nearest = torch.argmin(d2, dim=2)
        infl = infl * F.one_hot(nearest, kp.shape[0]).transpose(1, 2)
    elif aggregation_mode != "sum":
        raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")
    xpad = torch.cat([xf, torch.zeros_like(xf[:1])], 0)
    nx = xpad[idx]                                                # [n, H, Cin]
    wf = torch.matmul(infl, nx)                                   # [n, K, Cin]
    out = torch.matmul(wf.permute(1, 0, 2), w).sum(0)             # [n, Cout]
    num = (nx.sum(-1) > 0).sum(-1).clamp(min=1)
    return out / num[:, None].to(dtype)


def max_pool(x, inds) -> torch.Tensor:
    """max_pool (finegrained_kpconv_blocks.py:125-141): the zero shadow row takes part in the max."""
    x = torch.as_tensor(x).cpu()
    xpad = torch.cat([x, torch.zeros_like(x[:1])], 0)
    return xpad[torch.as_tensor(inds).long().cpu()].max(1)[0]


def _instance_norm(x: torch.Tensor, lens) -> torch.Tensor:
    """BatchNormBlock with nn.InstanceNorm1d (finegrained_kpconv_blocks.py:462-518): per cloud,
    per channel, biased variance, eps 1e-5, no affine, no running stats."""
    outs, i0 = [], 0
    for n in [int(v) for v in lens]:
        seg = x[i0:i0 + n]
        mu = seg.mean(0, keepdim=True)
        var = seg.var(0, unbiased=False, keepdim=True)
        outs.append((seg - mu) / torch.sqrt(var + 1e-5))
        i0 += n
    return torch.cat(outs, 0)


def _bn_eval(x, sd, prefix):
    """nn.BatchNorm1d in eval mode (running statistics), as inside my_Bottle2neck (res2net.py:100-121)."""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, 1e-5)


def _bn_train(x, sd, prefix):
    return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.0, 1e-5)


def _unary(x, lens, sd, prefix, relu=True):
    """UnaryBlock (finegrained_kpconv_blocks.py:521-555): Linear(no bias) -> InstanceNorm -> LeakyReLU(0.1)."""
    x = _instance_norm(x @ sd[prefix + ".mlp.weight"].t(), lens)
    return F.leaky_relu(x, 0.1) if relu else x


def _res2net(x, sd, prefix, training=False):
    """my_res2Net(my_Bottle2neck, C/4, C, baseWidth=14, scale=8) (res2net.py:84-159, 231-265)."""
    bn = _bn_train if training else _bn_eval
    p = prefix + ".layer1.0"
    out = F.relu(bn(x @ sd[p + ".conv1.weight"].t(), sd, p + ".bn1"))
    width = sd[p + ".convs.0.weight"].shape[0]
    splits = torch.split(out, width, 1)
    pieces, sp = [], None
    n_chain = len(splits) - 1
    for i in range(n_chain):
        sp = splits[i] if i == 0 else sp + splits[i]
        sp = F.relu(bn(sp @ sd[f"{p}.convs.{i}.weight"].t(), sd, f"{p}.bns.{i}"))
        pieces.append(sp)
    pieces.append(splits[n_chain])
    out = bn(torch.cat(pieces, 1) @ sd[p + ".conv3.weight"].t(), sd, p + ".bn3")
    if (p + ".downsample.0.weight") in sd:
        res = bn(x @ sd[p + ".downsample.0.weight"].t(), sd, p + ".downsample.1")
    else:
        res = x
    return F.relu(out + res)


def encoder_forward(sd: Dict[str, torch.Tensor], cfg, x, batch, training: bool = False, keep_grad: bool = False,
                    dtype=torch.float32):
    """KPFEncoder.forward (models/backbone_kpconv/finegrained_kpconv.py:86-95) over SimpleBlock /
    ResnetBottleneckBlock (finegrained_kpconv_blocks.py:578-634, 637-727), driven by a state_dict
    with the reference's parameter names (``encoder_blocks.{i}.KPConv.weights`` ...).
    Returns (features, skip_x)."""
    if not keep_grad:
        sd = {k: (v.detach().cpu().to(dtype) if v.is_floating_point() else v.detach().cpu()) for k, v in sd.items()}
    x = torch.as_tensor(x).to(dtype).cpu()
    pts = [torch.as_tensor(p).to(dtype).cpu() for p in batch["points"]]
    lens = [np.asarray(torch.as_tensor(l).cpu()) for l in batch["stack_lengths"]]
    r = cfg.first_subsampling_dl * cfg.conv_radius
    layer, skips = 0, []
    arch = list(cfg.architecture)
    # KPFEncoder.__init__ (:79-84): without a decoder the last block index is a skip index too
    last_skip = len(arch) - 1 if "upsample" not in arch[-1] else -1
    for bi, name in enumerate(arch):
        if "upsample" in name:
            break
        if any(t in name for t in ("pool", "strided", "upsample", "global")) or bi == last_skip:
            skips.append(x)
        strided = "strided" in name
        pre = f"encoder_blocks.{bi}"
        extent = r * cfg.KP_extent / cfg.conv_radius
        if strided:
            qp, sp, idx, post = pts[layer + 1], pts[layer], batch["pools"][layer], lens[layer + 1]
        else:
            qp, sp, idx, post = pts[layer], pts[layer], batch["neighbors"][layer], lens[layer]
        kw = dict(KP_extent=extent, KP_influence=cfg.KP_influence, aggregation_mode=cfg.aggregation_mode, dtype=dtype)
        if name.startswith("simple"):
            y = kpconv_forward(qp, sp, idx, x, sd[pre + ".KPConv.weights"], sd[pre + ".KPConv.kernel_points"], **kw)
            x = F.leaky_relu(_instance_norm(y, post), 0.1)
        elif name.startswith("resnetb"):
            feats = x
            y = _unary(feats, lens[layer], sd, pre + ".unary1") if (pre + ".unary1.mlp.weight") in sd else feats
            y = kpconv_forward(qp, sp, idx, y, sd[pre + ".KPConv.weights"], sd[pre + ".KPConv.kernel_points"], **kw)
            y = _instance_norm(y, post)
            y = F.leaky_relu(_res2net(y, sd, pre + ".res2net", training), 0.1)
            sc = max_pool(feats, idx) if strided else feats
            if (pre + ".unary_shortcut.mlp.weight") in sd:
                sc = _unary(sc, post, sd, pre + ".unary_shortcut", relu=False)
            x = F.leaky_relu(y + sc, 0.1)
        else:
            raise ValueError("Unknown block name in the architecture definition : " + name)
        if "pool" in name or strided:
            layer += 1
            r *= 2
    return x, skips


# --------------------------------------------------------------------------------------------
# Weighted Kabsch
# --------------------------------------------------------------------------------------------

def compute_overlaps(src_overlap, tgt_overlap, pools, level_sizes) -> List[np.ndarray]:
    """Per-level ground-truth overlap ratios (reference models/backbone_kpconv/finegrained_kpconv.py:545-571):
    level 0 = the per-point masks as float; level p = mean of level p-1 over the valid entries of pools[p-1]
    (index < row count of level p-1), clamped to [0, 1].  A row without a valid entry is 0/0 = nan, as in the reference."""
    level = np.concatenate([np.asarray(src_overlap), np.asarray(tgt_overlap)]).astype(np.float32)
    out = [level]
    for p in range(1, len(level_sizes)):
        idx = np.asarray(pools[p - 1]).astype(np.int64)
        valid = idx < int(level_sizes[p - 1])
        gathered = level[np.where(valid, idx, 0)] * valid
        with np.errstate(invalid="ignore", divide="ignore"):
            level = np.clip(gathered.sum(1, dtype=np.float32) / valid.sum(1).astype(np.float32), 0.0, 1.0).astype(np.float32)
        out.append(level)
    return out


def compute_rigid_transform(a, b, weights=None) -> torch.Tensor:
    """compute_rigid_transform (utils/se3_torch.py:131-173), torch CPU fp32, LAPACK SVD."""
    a = torch.as_tensor(a).float().cpu()
    b = torch.as_tensor(b).float().cpu()
    assert a.shape == b.shape and a.shape[-1] == 3
    if weights is not None:
        w = torch.as_tensor(weights).float().cpu()
        assert a.shape[:-1] == w.shape
        assert w.min() >= 0 and w.max() <= 1
        wn = w[..., None] / torch.clamp_min(w.sum(-1, keepdim=True)[..., None], 1e-6)
        ca = (a * wn).sum(-2)
        cb = (b * wn).sum(-2)
        cov = (a - ca[..., None, :]).transpose(-2, -1) @ ((b - cb[..., None, :]) * wn)
    else:
        ca, cb = a.mean(-2), b.mean(-2)
        cov = (a - ca[..., None, :]).transpose(-2, -1) @ (b - cb[..., None, :])
    u, _, vh = torch.linalg.svd(cov, full_matrices=True)
    v = vh.transpose(-2, -1)
    r_pos = v @ u.transpose(-1, -2)
    v_neg = v.clone()
    v_neg[..., 2] *= -1
    r_neg = v_neg @ u.transpose(-1, -2)
    rot = torch.where(torch.det(r_pos)[..., None, None] > 0, r_pos, r_neg)
    t = -rot @ ca[..., :, None] + cb[..., :, None]
    return torch.cat([rot, t], -1)


def fast_compute_rigid_transform(a, b, weights, weights_threshold: float = 0.85) -> torch.Tensor:
    """fast_compute_rigid_transform (utils/se3_torch.py:226-274): zero the weights that are not
    above the threshold (:240-242; the reference does it in place on a CUDA tensor), then the same
    Kabsch solve."""
    w = torch.as_tensor(weights).float().cpu()
    w = torch.where(w > weights_threshold, w, torch.zeros_like(w))
    return compute_rigid_transform(a, b, w)


def se3_compare(a, b):
    """se3_compare (utils/se3_torch.py:117-129): rotation error in degrees and translation error of a∘b⁻¹."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    rb_t = b[..., :3, :3].transpose(-1, -2)
    rot = a[..., :3, :3] @ rb_t
    trans = a[..., :3, 3] - (rot @ b[..., :3, 3:4])[..., 0]
    tr = rot[..., 0, 0] + rot[..., 1, 1] + rot[..., 2, 2]
    deg = torch.acos(torch.clamp(0.5 * (tr - 1), -1.0, 1.0)) * 180.0 / math.pi
    return {"rot_deg": deg, "trans": trans.norm(dim=-1)}


def pose_error(a, b):
    """Well-conditioned pose difference for the parity tolerances (1e-3 deg, 1e-5 m): the rotation
    angle of Ra Rb^T from the Frobenius norm, theta = 2 asin(|Ra Rb^T - I|_F / (2 sqrt 2)) (the
    acos-of-trace form of se3_compare cannot resolve angles below ~0.03 deg in fp32), and |ta - tb|."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    rel = a[..., :3, :3] @ b[..., :3, :3].transpose(-1, -2)
    eye = torch.eye(3, dtype=torch.float64).expand_as(rel)
    fro = (rel - eye).flatten(-2).norm(dim=-1)
    deg = 2.0 * torch.asin(torch.clamp(fro / (2.0 * math.sqrt(2.0)), max=1.0)) * 180.0 / math.pi
    return {"rot_deg": deg, "trans": (a[..., :3, 3] - b[..., :3, 3]).norm(dim=-1)}
