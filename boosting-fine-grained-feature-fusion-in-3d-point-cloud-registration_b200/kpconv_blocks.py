"""KPConv operator and encoder blocks — host-side mirror of the reference's
``models/backbone_kpconv/finegrained_kpconv_blocks.py`` with the irregular work on CUDA kernels.

Same class names, constructor arguments, attribute / parameter names (so reference checkpoints load
with ``strict=True``) and call signatures:

* ``KPConv(kernel_size, p_dim, in_channels, out_channels, KP_extent, radius, ...)`` with parameters
  ``weights [K,Cin,Cout]`` and ``kernel_points [K,3]`` (reference :171-263); ``forward(q_pts, s_pts,
  neighb_inds, x)`` (:265-401) runs ``kpreg_kpconv_forward`` / ``kpreg_kpconv_backward``.
* ``gather`` (:66-97), ``max_pool`` (:125-141, CUDA), ``closest_pool`` (:110-122), ``global_average`` (:144-163).
* ``BatchNormBlock`` (:462-518, per-cloud InstanceNorm, here without the Python loop over clouds),
  ``UnaryBlock`` (:521-555), ``SimpleBlock`` (:578-634), ``ResnetBottleneckBlock`` (:637-727),
  pool / upsample blocks (:729-771) and ``block_decider`` (:414-460).

Deformable KPConv is not part of the registration path (no shipped config uses it) and raises.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from . import ops
from .kernel_points import load_kernels
from .res2net import my_Bottle2neck, my_res2Net

# contraction back end used by KPConv.forward: 1 = tcgen05 (3xTF32) tensor-core GEMM, 0 = fp32 CUDA cores
DEFAULT_GEMM = 1
# inference (no autograd) on CUDA: run the blocks' Linear / InstanceNorm / eval-BatchNorm / activation glue on
# the fused CUDA kernels (kpreg_linear_forward, kpreg_segment_norm_forward); otherwise stock PyTorch ops
FUSED_GLUE = True
# res2net's chained layers in one register-resident kernel (kpreg_chain_forward) where the width allows it
CHAIN_KERNEL = True
# conv1 + the chained layers in one tcgen05 kernel (kpreg_front_forward) where width / input channels allow it
# (KPREG_NO_FRONT=1: conv1 as its own GEMM + the chain kernel, for A/B runs)
import os as _os
# (the front kernel always splits operands into fp16 pairs: under KPREG_GEMM_TF32=1 — fp32 range — it is not used)
FRONT_KERNEL = _os.environ.get("KPREG_NO_FRONT", "")[:1] != "1" and _os.environ.get("KPREG_GEMM_TF32", "")[:1] != "1"
SHARED_UNARY = _os.environ.get("KPREG_NO_SHARED_UNARY", "")[:1] != "1"  # unary1 and the shortcut's unary of a block as ONE GEMM over the block input
PAIR_CONV3 = _os.environ.get("KPREG_NO_PAIR_CONV3", "")[:1] != "1"  # wide res2net units: conv3 + residual projection over (z, x) without copying x


def _fused(x: torch.Tensor) -> bool:
    return FUSED_GLUE and x.is_cuda and not torch.is_grad_enabled()


def gather(x, idx, method=2):
    """x[idx] for x [N, D...] and idx [n_1..n_m] -> [n_1..n_m, D...] (all three reference methods are
    the same indexing; they differ only in autograd cost on the reference's eager path)."""
    if method not in (0, 1, 2):
        raise ValueError("Unkown method")
    return x[idx.long()]


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inds, order=None):
        out, arg = ops.max_pool_forward(x, inds, want_argmax=ctx.needs_input_grad[0], order=order)
        ctx.n_s = x.shape[0]
        if arg is not None:
            ctx.save_for_backward(arg)
        return out

    @staticmethod
    def backward(ctx, grad):
        (arg,) = ctx.saved_tensors
        return ops.max_pool_backward(grad.contiguous(), arg, ctx.n_s), None, None


def max_pool(x, inds, order=None):
    """[n2, d] = max over the pooling rows of [x; 0] (the zero shadow row takes part).
    ``order`` (optional, not in the reference): processing-order permutation of the pooled rows."""
    return _MaxPoolFn.apply(x, inds, order)


def closest_pool(x, inds):
    """Features of the closest (first-column) neighbour; shadow index -> zeros."""
    x = torch.cat((x, torch.zeros_like(x[:1, :])), 0)
    return x[inds[:, 0].long()]


def global_average(x, batch_lengths):
    """[B, D] per-cloud mean of x [N, D]."""
    lens = batch_lengths.to(device=x.device, dtype=torch.long)
    seg = torch.repeat_interleave(torch.arange(lens.shape[0], device=x.device), lens, output_size=x.shape[0])
    sums = torch.zeros((lens.shape[0], x.shape[1]), dtype=x.dtype, device=x.device).index_add_(0, seg, x)
    return sums / lens[:, None].to(x.dtype)


class _KPConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_pts, s_pts, neighb_inds, x, weights, kernel_points, extent, influence, aggregation, gemm, order):
        ctx.save_for_backward(q_pts, s_pts, neighb_inds, x, weights, kernel_points)
        ctx.cfg = (extent, influence, aggregation)
        ctx.order = order
        return ops.kpconv_forward(q_pts, s_pts, neighb_inds, x, weights, kernel_points, extent, influence,
                                  aggregation, gemm, order)

    @staticmethod
    def backward(ctx, grad_out):
        q_pts, s_pts, neighb_inds, x, weights, kernel_points = ctx.saved_tensors
        extent, influence, aggregation = ctx.cfg
        d_x, d_w = ops.kpconv_backward(q_pts, s_pts, neighb_inds, x, weights, kernel_points, grad_out.contiguous(),
                                       extent, influence, aggregation, ctx.order)
        # coordinates, indices and kernel points carry no gradient (reference :262-263)
        return None, None, None, d_x, d_w, None, None, None, None, None, None


class KPConv(ops.CacheInvalidatingModule):

    def __init__(self, kernel_size, p_dim, in_channels, out_channels, KP_extent, radius,
                 fixed_kernel_points='center', KP_influence='linear', aggregation_mode='sum',
                 deformable=False, modulated=False):
        super().__init__()
        if deformable:
            raise NotImplementedError("deformable KPConv is outside the registration hot path")
        if KP_influence not in ops.INFLUENCE:
            raise ValueError('Unknown influence function type (config.KP_influence)')
        if aggregation_mode not in ops.AGGREGATION:
            raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")
        self.K = kernel_size
        self.p_dim = p_dim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.radius = radius
        self.KP_extent = KP_extent
        self.fixed_kernel_points = fixed_kernel_points
        self.KP_influence = KP_influence
        self.aggregation_mode = aggregation_mode
        self.deformable = deformable
        self.modulated = modulated
        self.gemm = None  # None -> module-level DEFAULT_GEMM

        self.weights = Parameter(torch.zeros((self.K, in_channels, out_channels), dtype=torch.float32),
                                 requires_grad=True)
        self.reset_parameters()
        pts = load_kernels(self.radius, self.K, dimension=self.p_dim, fixed=self.fixed_kernel_points)
        self.kernel_points = Parameter(torch.tensor(pts, dtype=torch.float32), requires_grad=False)

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))

    def forward(self, q_pts, s_pts, neighb_inds, x, order=None):
        """``order`` (optional, not in the reference): int32 permutation of the query rows giving a spatially
        coherent processing order (Preprocessor's batch['orders']); it never changes the result."""
        gemm = DEFAULT_GEMM if self.gemm is None else self.gemm
        if not torch.is_grad_enabled():
            # inference: no autograd node (x reaches ops.kpconv_forward as the very tensor the norm kernel tagged with its row predicate)
            return ops.kpconv_forward(q_pts, s_pts, neighb_inds, x, self.weights, self.kernel_points, float(self.KP_extent),
                                      self.KP_influence, self.aggregation_mode, gemm, order)
        return _KPConvFn.apply(q_pts, s_pts, neighb_inds, x, self.weights, self.kernel_points, float(self.KP_extent),
                               self.KP_influence, self.aggregation_mode, gemm, order)

    def __repr__(self):
        return 'KPConv(radius: {:.2f}, extent: {:.2f}, in_feat: {:d}, out_feat: {:d})'.format(
            self.radius, self.KP_extent, self.in_channels, self.out_channels)


_NORM_LANES = 64  # sub-accumulators per cloud: spreads the atomics of index_add_ over 64 rows


def _segment_sums(x: torch.Tensor, seg: torch.Tensor, n_seg: int) -> torch.Tensor:
    """Per-segment column sums of x [N,C] -> [n_seg,C] (fp32, two-level to keep atomic contention low)."""
    lane = torch.arange(x.shape[0], device=x.device) % _NORM_LANES
    part = torch.zeros((n_seg * _NORM_LANES, x.shape[1]), dtype=x.dtype, device=x.device)
    part.index_add_(0, seg * _NORM_LANES + lane, x)
    return part.view(n_seg, _NORM_LANES, x.shape[1]).sum(1)


def _segment_instance_norm(x: torch.Tensor, stack_lengths: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """InstanceNorm1d (affine=False, no running stats) of every cloud of a stacked [N,C] tensor:
    per cloud and channel, (x - mean) / sqrt(biased var + eps).  Two-pass, no host sync."""
    n, _ = x.shape
    lens = stack_lengths.to(device=x.device, dtype=torch.long)
    n_seg = lens.shape[0]
    seg = torch.repeat_interleave(torch.arange(n_seg, device=x.device), lens, output_size=n)
    cnt = lens.clamp(min=1).to(x.dtype)[:, None]
    mean = _segment_sums(x, seg, n_seg) / cnt
    cen = x - mean[seg]
    var = _segment_sums(cen * cen, seg, n_seg) / cnt
    return cen * torch.rsqrt(var + eps)[seg]


class _SegNormFn(torch.autograd.Function):
    """Per-cloud instance norm with both directions on the CUDA kernels (training path: the torch formulation's
    ``mean[seg]`` gathers back-propagate through index_put_ with accumulation, 60 % of an 8-pair training step)."""

    @staticmethod
    def forward(ctx, x, stack_lengths):
        ctx.save_for_backward(x, stack_lengths)
        return ops.segment_norm(x, stack_lengths)

    @staticmethod
    def backward(ctx, grad):
        x, stack_lengths = ctx.saved_tensors
        return ops.segment_norm_backward(x, grad.contiguous(), stack_lengths), None


class BatchNormBlock(nn.Module):

    def __init__(self, in_dim, use_bn, bn_momentum):
        super().__init__()
        self.bn_momentum = bn_momentum
        self.use_bn = use_bn
        self.in_dim = in_dim
        if not self.use_bn:
            self.bias = Parameter(torch.zeros(in_dim, dtype=torch.float32), requires_grad=True)

    def reset_parameters(self):
        nn.init.zeros_(self.bias)

    def forward(self, x, stack_lengths, act=None, residual=None, row_pos=False):
        """``act`` / ``residual`` let the fused path apply the activation (and shortcut addition) that
        follows the norm in the same kernel; ``row_pos``: the result feeds a KPConv (ops.segment_norm); the default
        arguments are the reference's signature."""
        if self.use_bn:
            if _fused(x) and x.shape[1] % 4 == 0:
                return ops.segment_norm(x, stack_lengths, residual=residual, act=act, slope=0.1, row_pos=row_pos)
            if x.is_cuda and x.shape[1] % 4 == 0 and x.shape[0] > 0:
                x = _SegNormFn.apply(x, stack_lengths.to(x.device))
            else:
                x = _segment_instance_norm(x, stack_lengths)
        else:
            x = x + self.bias
        if residual is not None:
            x = x + residual
        return F.leaky_relu(x, 0.1) if act == "leaky_relu" else x

    def __repr__(self):
        return 'BatchNormBlock(in_feat: {:d}, momentum: {:.3f}, only_bias: {:s})'.format(
            self.in_dim, self.bn_momentum, str(not self.use_bn))


class UnaryBlock(ops.CacheInvalidatingModule):

    def __init__(self, in_dim, out_dim, use_bn, bn_momentum, no_relu=False):
        super().__init__()
        self.bn_momentum = bn_momentum
        self.use_bn = use_bn
        self.no_relu = no_relu
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.mlp = nn.Linear(in_dim, out_dim, bias=False)
        self.batch_norm = BatchNormBlock(out_dim, self.use_bn, self.bn_momentum)
        if not no_relu:
            self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, x, stack_lengths=None, residual=None, final_act=False, feeds_kpconv=False):
        """residual / final_act (fused inference only): leaky_relu(norm(mlp(x)) + residual); feeds_kpconv: the result is the
        feature input of a KPConv (its row predicate is then written by the norm kernel)."""
        if _fused(x):
            y = ops.linear_forward(x, self.mlp.weight, gemm=DEFAULT_GEMM)
            act = "leaky_relu" if (final_act or not self.no_relu) else None
            return self.batch_norm(y, stack_lengths, act=act, residual=residual, row_pos=feeds_kpconv)
        x = self.batch_norm(ops.linear_train(x, self.mlp), stack_lengths)
        x = x if self.no_relu else self.leaky_relu(x)
        if residual is not None:
            x = x + residual
        return F.leaky_relu(x, 0.1) if final_act else x

    def __repr__(self):
        return 'UnaryBlock(in_feat: {:d}, out_feat: {:d}, BN: {:s}, ReLU: {:s})'.format(
            self.in_dim, self.out_dim, str(self.use_bn), str(not self.no_relu))


class UnaryBlock2(nn.Module):
    """Linear - ReLU - Linear."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(in_dim, in_dim), nn.ReLU(), nn.Linear(in_dim, out_dim))
        self.in_dim = in_dim
        self.out_dim = out_dim

    def forward(self, x):
        return self.mlp(x)


def _conv_inputs(batch, layer_ind, strided):
    """(q_pts, s_pts, neighb_inds, stack_lengths of the output level, processing order of the query rows or None)
    for a block at pyramid level layer_ind."""
    orders = batch.get('orders') if hasattr(batch, 'get') else None
    q_level = layer_ind + 1 if strided else layer_ind
    order = orders[q_level] if orders is not None and q_level < len(orders) else None
    if order is not None and not order.is_cuda:
        order = None
    if strided:
        return (batch['points'][layer_ind + 1], batch['points'][layer_ind], batch['pools'][layer_ind],
                batch['stack_lengths'][layer_ind + 1], order)
    return (batch['points'][layer_ind], batch['points'][layer_ind], batch['neighbors'][layer_ind],
            batch['stack_lengths'][layer_ind], order)


def _make_kpconv(in_dim, out_dim, radius, config, block_name):
    return KPConv(config.num_kernel_points, config.in_points_dim, in_dim, out_dim,
                  radius * config.KP_extent / config.conv_radius, radius,
                  fixed_kernel_points=config.fixed_kernel_points, KP_influence=config.KP_influence,
                  aggregation_mode=config.aggregation_mode, deformable='deform' in block_name,
                  modulated=config.modulated)


class SimpleBlock(nn.Module):
    """KPConv(in -> out/2) -> per-cloud norm -> LeakyReLU(0.1)."""

    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config):
        super().__init__()
        self.bn_momentum = config.batch_norm_momentum
        self.use_bn = config.use_batch_norm
        self.layer_ind = layer_ind
        self.block_name = block_name
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.KPConv = _make_kpconv(in_dim, out_dim // 2, radius, config, block_name)
        self.batch_norm = BatchNormBlock(out_dim // 2, self.use_bn, self.bn_momentum)
        self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, x, batch):
        q_pts, s_pts, inds, lens, order = _conv_inputs(batch, self.layer_ind, 'strided' in self.block_name)
        return self.batch_norm(self.KPConv(q_pts, s_pts, inds, x, order), lens, act="leaky_relu")


class ResnetBottleneckBlock(nn.Module):
    """unary(in -> C/4) -> KPConv(C/4 -> C/4) -> norm -> res2net(C/4 -> C) -> LeakyReLU, plus the
    (max-pooled when strided) shortcut through an optional unary, summed and activated."""

    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config, flag=False):
        super().__init__()
        self.bn_momentum = config.batch_norm_momentum
        self.use_bn = config.use_batch_norm
        self.block_name = block_name
        self.layer_ind = layer_ind
        self.in_dim = in_dim
        self.out_dim = out_dim
        mid = out_dim // 4
        self.unary1 = UnaryBlock(in_dim, mid, self.use_bn, self.bn_momentum) if in_dim != mid else nn.Identity()
        self.KPConv = _make_kpconv(mid, mid, radius, config, block_name)
        self.batch_norm_conv = BatchNormBlock(mid, self.use_bn, self.bn_momentum)
        self.res2net = my_res2Net(my_Bottle2neck, mid, out_dim, baseWidth=14, scale=8)
        self.unary_shortcut = (UnaryBlock(in_dim, out_dim, self.use_bn, self.bn_momentum, no_relu=True)
                               if in_dim != out_dim else nn.Identity())
        self.leaky_relu = nn.LeakyReLU(0.1)

    def _joint_unary_weight(self):
        """[W_unary1 ; W_shortcut] ([mid + out, in]): both Linear layers read the block input, one GEMM reads it once."""
        w1, w2 = self.unary1.mlp.weight, self.unary_shortcut.mlp.weight
        key = (w1._version, w2._version, w1.data_ptr(), w2.data_ptr(), ops.cache_epoch())
        cache = getattr(self, "_kpreg_joint_unary", None)
        if cache is None or cache[0] != key:
            with torch.no_grad():
                cache = (key, torch.cat([w1, w2], 0).contiguous())
            self._kpreg_joint_unary = cache
        return cache[1]

    def forward(self, features, batch):
        strided = 'strided' in self.block_name
        q_pts, s_pts, inds, lens_post, order = _conv_inputs(batch, self.layer_ind, strided)
        lens_pre = batch['stack_lengths'][self.layer_ind]

        if (SHARED_UNARY and not strided and _fused(features) and self.use_bn and isinstance(self.unary1, UnaryBlock)
                and isinstance(self.unary_shortcut, UnaryBlock) and self.unary1.out_dim % 4 == 0 and features.shape[0] > 0):
            # inference, shortcut through a unary: y = features [W1 ; Ws]^T in one GEMM (the block input crosses HBM once), then
            # the two norms on column slices of y — the arithmetic of every output column is that of the separate layers
            mid = self.unary1.out_dim
            y = ops.linear_forward(features, self._joint_unary_weight(), gemm=DEFAULT_GEMM)
            x = self.unary1.batch_norm(y[:, :mid], lens_pre, act="leaky_relu", row_pos=True)
            x = self.batch_norm_conv(self.KPConv(q_pts, s_pts, inds, x, order), lens_post)
            x = self.res2net(x)
            return self.unary_shortcut.batch_norm(y[:, mid:], lens_post, act="leaky_relu", residual=x)

        x = self.unary1(features, lens_pre, feeds_kpconv=True) if isinstance(self.unary1, UnaryBlock) else features
        x = self.batch_norm_conv(self.KPConv(q_pts, s_pts, inds, x, order), lens_post)
        if _fused(x) and not isinstance(self.unary_shortcut, UnaryBlock):
            # identity shortcut: leaky_relu(res2net(x) + shortcut) rides on res2net's last GEMM
            return self.res2net(x, max_pool(features, inds, order) if strided else features)
        x = self.res2net(x)
        if not _fused(x):
            # my_Bottle2neck ends in a ReLU, so this LeakyReLU is the identity; kept on the autograd path only so
            # that the graph matches the reference op for op
            x = self.leaky_relu(x)

        shortcut = max_pool(features, inds, order) if strided else features
        if isinstance(self.unary_shortcut, UnaryBlock):
            # leaky_relu(x + norm(mlp(shortcut))): the addition and activation ride on the norm kernel when fused
            return self.unary_shortcut(shortcut, lens_post, residual=x, final_act=True)
        return self.leaky_relu(x + shortcut)


class GlobalAverageBlock(nn.Module):
    def forward(self, x, batch):
        return global_average(x, batch['stack_lengths'][-1])


class NearestUpsampleBlock(nn.Module):
    def __init__(self, layer_ind):
        super().__init__()
        self.layer_ind = layer_ind

    def forward(self, x, batch):
        return closest_pool(x, batch['upsamples'][self.layer_ind - 1])

    def __repr__(self):
        return 'NearestUpsampleBlock(layer: {:d} -> {:d})'.format(self.layer_ind, self.layer_ind - 1)


class MaxPoolBlock(nn.Module):
    def __init__(self, layer_ind):
        super().__init__()
        self.layer_ind = layer_ind

    def forward(self, x, batch):
        return max_pool(x, batch['pools'][self.layer_ind + 1])


_SIMPLE = {'simple' + a + b for a in ('', '_deformable', '_invariant', '_equivariant') for b in ('', '_strided')}
_RESNET = {'resnetb' + a + b for a in ('', '_deformable', '_invariant', '_equivariant') for b in ('', '_strided')}


def block_decider(block_name, radius, in_dim, out_dim, layer_ind, config, flag=False):
    if block_name == 'unary':
        return UnaryBlock(in_dim, out_dim, config.use_batch_norm, config.batch_norm_momentum)
    if block_name == 'unary2':
        return UnaryBlock2(in_dim, out_dim)
    if block_name in _SIMPLE:
        return SimpleBlock(block_name, in_dim, out_dim, radius, layer_ind, config)
    if block_name in _RESNET:
        return ResnetBottleneckBlock(block_name, in_dim, out_dim, radius, layer_ind, config, flag)
    if block_name in ('max_pool', 'max_pool_wide'):
        return MaxPoolBlock(layer_ind)
    if block_name == 'global_average':
        return GlobalAverageBlock()
    if block_name == 'nearest_upsample':
        return NearestUpsampleBlock(layer_ind)
    raise ValueError('Unknown block name in the architecture definition : ' + block_name)
