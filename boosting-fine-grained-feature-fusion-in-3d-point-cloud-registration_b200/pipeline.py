"""The registration hot path end to end, batched and pair-sharded.

One step = for a batch of point-cloud pairs: KPConv pyramid (``Preprocessor``) -> ``KPFEncoder``
forward -> weighted Kabsch pose per pair.  This is the slice of ``RegTR.forward`` that BASELINE.json
names (reference models/finegrained_regtr.py:121, :139, :215-218).  The transformer / correspondence
decoder that sits between the encoder and the pose solve in the reference is out of scope, so the
correspondences fed to the Kabsch stage are synthetic but shaped like the decoder's output
(``[6, N_src_c + N_tgt_c, 3]`` per pair on the coarsest pyramid level).

Multi-GPU: pairs are independent, so they are sharded round-robin over ranks (one process per GPU) with
no data-path collective; the only exchange is an all-gather of the per-pair ``[3,4]`` poses and the
``(rot_deg, trans)`` errors (NCCL over NVLink; gloo on CPU for the tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import os

import torch
import torch.distributed as dist

from . import ops
from .kpconv import KPFEncoder, Preprocessor
from .se3_torch import compute_rigid_transform_batch, se3_compare

N_DECODER_LAYERS = 6  # RegTR predicts a pose after each of its 6 decoder layers


def shard_pairs(n_pairs: int, rank: int, world_size: int) -> List[int]:
    """Pair i is processed by rank i mod world_size."""
    return list(range(rank, n_pairs, world_size))


def gather_results(local: torch.Tensor, n_pairs: int, rank: int, world_size: int) -> torch.Tensor:
    """All-gather per-pair result rows.  ``local`` is [n_local, D] for the pairs of shard_pairs(); every
    rank returns the full [n_pairs, D] table in pair order.  Shards are padded to equal length."""
    if world_size == 1:
        return local
    per = (n_pairs + world_size - 1) // world_size
    padded = torch.zeros((per, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world_size * per, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    # row (r, j) of the gathered table is pair r + j*world_size
    table = out.view(world_size, per, -1).transpose(0, 1).reshape(world_size * per, -1)
    return table[:n_pairs]


def synthetic_correspondences(coarse_pts: torch.Tensor, coarse_lens: Sequence[int], poses: torch.Tensor,
                              noise: float = 0.01, seed: int = 0):
    """Decoder-shaped correspondences for every pair of a stacked batch (first B clouds = sources, last B =
    targets, the reference's ``split_src_tgt`` convention, utils/seq_manipulation.py:42-48).

    For pair p: a = [src_c ; R^-1 (tgt_c - t) + n], b = [R src_c + t + n ; tgt_c], repeated for the 6
    decoder layers with fresh noise, weights = sigmoid(N(0, 2)).  Returns lists (a, b, w) of
    [6, Nc_p, 3] / [6, Nc_p] tensors on the device of ``coarse_pts``."""
    n_pairs = len(coarse_lens) // 2
    starts = [0]
    for n in coarse_lens:
        starts.append(starts[-1] + int(n))
    gen = torch.Generator(device=coarse_pts.device)
    gen.manual_seed(seed)
    a_all, b_all, w_all = [], [], []
    for p in range(n_pairs):
        src = coarse_pts[starts[p]:starts[p + 1]]
        tgt = coarse_pts[starts[n_pairs + p]:starts[n_pairs + p + 1]]
        rot, trans = poses[p, :, :3], poses[p, :, 3]
        n_c = src.shape[0] + tgt.shape[0]
        a = torch.cat([src, (tgt - trans) @ rot], 0).expand(N_DECODER_LAYERS, n_c, 3)
        b = torch.cat([src @ rot.T + trans, tgt], 0).expand(N_DECODER_LAYERS, n_c, 3)
        jitter = noise * torch.randn((2, N_DECODER_LAYERS, n_c, 3), generator=gen, device=coarse_pts.device)
        mask = torch.zeros((n_c, 1), device=coarse_pts.device)
        mask[src.shape[0]:] = 1.0
        a_all.append((a + jitter[0] * mask).contiguous())
        b_all.append((b + jitter[1] * (1.0 - mask)).contiguous())
        w_all.append(torch.sigmoid(2.0 * torch.randn((N_DECODER_LAYERS, n_c), generator=gen, device=coarse_pts.device)))
    return a_all, b_all, w_all


def synthetic_correspondences_batched(coarse_pts: torch.Tensor, coarse_lens: torch.Tensor, poses: torch.Tensor,
                                      noise: float = 0.01, seed: int = 0):
    """Same construction as synthetic_correspondences for ALL pairs at once, entirely on the device and without
    a host read-back (a handful of launches whatever the batch size).  Returns (a, b [T,3], w [T], offsets int64
    [6B+1]): set (pair p, layer l) covers rows offsets[6p+l] .. offsets[6p+l+1] — the ragged layout kpreg_kabsch takes.
    The random draws differ from the per-pair version (one generator stream for the whole batch)."""
    dev = coarse_pts.device
    n_c = coarse_pts.shape[0]
    n_clouds = coarse_lens.shape[0]
    n_pairs = n_clouds // 2
    lens = coarse_lens.to(torch.long)
    cloud = torch.repeat_interleave(torch.arange(n_clouds, device=dev), lens, output_size=n_c)
    pair = cloud % n_pairs
    is_tgt = (cloud >= n_pairs)
    # group the points of a pair together: sources first, then targets (stable sort keeps their order)
    perm = torch.argsort(pair, stable=True)
    pts, pair_s, tgt_s = coarse_pts[perm], pair[perm], is_tgt[perm][:, None]
    rot, trans = poses[pair_s, :, :3], poses[pair_s, :, 3]
    fwd = torch.einsum('nij,nj->ni', rot, pts) + trans               # R p + t
    back = torch.einsum('nji,nj->ni', rot, pts - trans)              # R^T (p - t)
    a0 = torch.where(tgt_s, back, pts)
    b0 = torch.where(tgt_s, pts, fwd)
    n_pair = lens[:n_pairs] + lens[n_pairs:]                          # points per pair
    cum = torch.nn.functional.pad(torch.cumsum(n_pair, 0), (1, 0))   # [B+1]
    # flat position f of (pair p, layer l, j): 6 * cum[p] + l * n_p + j
    f = torch.arange(N_DECODER_LAYERS * n_c, device=dev)
    p_of_f = torch.searchsorted(N_DECODER_LAYERS * cum[1:], f, right=True)
    j = (f - N_DECODER_LAYERS * cum[p_of_f]) % n_pair[p_of_f]
    src_row = cum[p_of_f] + j
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    jitter = noise * torch.randn((2, N_DECODER_LAYERS * n_c, 3), generator=gen, device=dev)
    t_rows = tgt_s[src_row]
    a = a0[src_row] + jitter[0] * t_rows
    b = b0[src_row] + jitter[1] * (~t_rows)
    w = torch.sigmoid(2.0 * torch.randn(N_DECODER_LAYERS * n_c, generator=gen, device=dev))
    layer = torch.arange(N_DECODER_LAYERS, device=dev)
    offsets = (N_DECODER_LAYERS * cum[:-1, None] + layer[None, :] * n_pair[:, None]).reshape(-1)
    offsets = torch.cat([offsets, (N_DECODER_LAYERS * cum[-1]).reshape(1)]).to(torch.int64)
    return a.contiguous(), b.contiguous(), w.contiguous(), offsets


class RegistrationPath(torch.nn.Module):
    """Preprocessor + KPFEncoder + batched Kabsch behind one call."""

    def __init__(self, cfg, index_dtype: torch.dtype = torch.int32, weights_threshold: Optional[float] = None,
                 overlap: Optional[bool] = None):
        super().__init__()
        self.cfg = cfg
        self.preprocessor = Preprocessor(cfg, index_dtype=index_dtype)
        self.kpf_encoder = KPFEncoder(cfg, cfg.d_embed)
        self.weights_threshold = weights_threshold
        # overlap (off by default; KPREG_OVERLAP=1 or overlap=True): build the pyramid below the finest level on a second CUDA
        # stream while the encoder's finest-level blocks run.  Results are identical — only the launch order across streams
        # changes — and so, measured on B200, is the step time (53.5-54 ms either way at 64 pairs): every kernel of the step
        # is launched wide enough to fill all SMs, so the two streams take turns instead of sharing them.
        self.overlap = (os.environ.get("KPREG_OVERLAP", "")[:1] == "1") if overlap is None else bool(overlap)
        self._side = {}

    def _pyramid_and_encoder(self, clouds):
        """(meta, encoder output).  Overlapped form: the radius queries / subsampling of the coarser levels are bound by
        instruction issue and touch almost no HBM, the finest level's Linear / norm layers are bound by HBM — so after the
        finest level's conv table the rest of the pyramid is issued on a side stream, behind the (already enqueued)
        finest-level encoder blocks of the main stream.  Stream safety: the side stream first waits for everything the main
        stream has been given (its inputs; tensors of earlier steps last read there), the main stream waits for the side
        stream before the first strided block, and no main-stream allocation happens while the side stream is being fed."""
        enc, pre = self.kpf_encoder, self.preprocessor
        dev = clouds[0].device
        first = enc.first_strided_block()
        if not (self.overlap and dev.type == "cuda" and 0 < first < len(enc.encoder_blocks)):
            meta = pre(clouds)
            feats0 = torch.ones((meta['points'][0].shape[0], 1), dtype=torch.float32, device=meta['points'][0].device)
            feats, _ = enc(feats0, meta)
            return meta, feats
        main = torch.cuda.current_stream(dev)
        side = self._side.get(dev)
        if side is None:
            side = self._side[dev] = torch.cuda.Stream(device=dev)
        stages = pre.stages(clouds, staged=True)
        meta0 = next(stages)
        feats0 = torch.ones((meta0['points'][0].shape[0], 1), dtype=torch.float32, device=dev)
        x, skips = enc.forward_blocks(feats0, meta0, 0, first, [])
        side.wait_stream(main)
        with torch.cuda.stream(side):
            meta = next(stages)
            for _ in stages:  # run the generator to its end
                pass
        main.wait_stream(side)
        feats, _ = enc.forward_blocks(x, meta, first, len(enc.encoder_blocks), skips)
        return meta, feats

    @torch.no_grad()
    def forward(self, src_xyz: Sequence[torch.Tensor], tgt_xyz: Sequence[torch.Tensor], poses_gt: torch.Tensor,
                corr_seed: int = 0, corr=None) -> Dict[str, torch.Tensor]:
        """src_xyz / tgt_xyz: lists of [N_i,3] clouds (the batch dict of collate_pair,
        data_loaders/collate_functions.py:13-22); poses_gt [B,3,4] seeds the synthetic correspondences.
        ``corr`` = (a, b, w, offsets) as returned in the result's 'corr' entry: decoder-shaped correspondences prepared
        by the caller (in the model they are the decoder's output; building the synthetic stand-ins is not part of the
        path).  Returns the encoder features, the per-layer poses [6,B,3,4] and the final-layer pose errors."""
        n_pairs = len(src_xyz)
        meta, feats = self._pyramid_and_encoder(list(src_xyz) + list(tgt_xyz))
        coarse = meta['points'][-1]
        poses_dev = poses_gt.to(coarse.device)
        if corr is not None:
            a, b, w, offsets = corr
            if int(a.shape[0]) != N_DECODER_LAYERS * int(coarse.shape[0]):
                raise RuntimeError("corr does not match this batch's coarse level")
        else:
            a, b, w, offsets = synthetic_correspondences_batched(coarse, meta['stack_lengths'][-1], poses_dev, seed=corr_seed)
        thr = -1.0 if self.weights_threshold is None else float(self.weights_threshold)
        poses = ops.kabsch(a, b, w, n_pairs * N_DECODER_LAYERS, 0, offsets, thr, False)
        poses = poses.reshape(n_pairs, N_DECODER_LAYERS, 3, 4).transpose(0, 1).contiguous()   # [6, B, 3, 4]
        err = se3_compare(poses[-1], poses_dev)
        return {'feats': feats, 'poses': poses, 'rot_deg': err['rot_deg'], 'trans': err['trans'], 'meta': meta,
                'n_pairs': n_pairs, 'corr': (a, b, w, offsets)}


def result_rows(out: Dict[str, torch.Tensor]) -> torch.Tensor:
    """[B, 14] rows: the final-layer [3,4] pose flattened + (rot_deg, trans) — what the ranks exchange."""
    pose = out['poses'][-1].reshape(-1, 12)
    return torch.cat([pose, out['rot_deg'][:, None], out['trans'][:, None]], 1).contiguous()
