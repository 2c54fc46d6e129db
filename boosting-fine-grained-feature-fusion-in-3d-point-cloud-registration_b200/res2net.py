"""The fork's "fine-grained feature fusion" unit: a Res2Net-style multi-scale MLP on point features.

Mirror of ``my_Bottle2neck`` / ``my_res2Net`` in the reference's ``models/backbone_kpconv/res2net.py``
(:84-159, :231-265) with identical sub-module names (``layer1.0.conv1``, ``bn1``, ``convs.i``, ``bns.i``,
``conv3``, ``bn3``, ``downsample.0/1``) so checkpoints are interchangeable.  In inference on CUDA (width a
multiple of 4) the unit runs on the fused path (``_fused_forward``): Linear + eval-BatchNorm (+ ReLU) as tcgen05
GEMMs with fused epilogues, the chained layers in one register-resident kernel; with autograd enabled or in training
mode the layers are stock PyTorch ops.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def _cache_base():
    from . import ops
    return ops.CacheInvalidatingModule


_CacheInvalidating = _cache_base()


def _folded(linear: nn.Linear, bn: nn.BatchNorm1d):
    """(weight', shift) with the eval-mode BatchNorm1d folded into the bias-free Linear that precedes it:
    bn(x W^T) = x (diag(s) W)^T + (beta - mean * s),  s = gamma / sqrt(var + eps).
    Cached on the BatchNorm module until a parameter or running statistic changes (version counters) or
    ``ops.invalidate_caches()`` is called (needed after writes through ``.data``, which bypass the counters)."""
    from . import ops
    key = (linear.weight._version, bn.weight._version, bn.bias._version, bn.running_mean._version,
           bn.running_var._version, linear.weight.data_ptr(), bn.running_mean.data_ptr(), ops.cache_epoch())
    cache = getattr(bn, "_kpreg_folded", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
            cache = (key, (linear.weight * scale[:, None]).contiguous(), (bn.bias - bn.running_mean * scale).contiguous())
        bn._kpreg_folded = cache
    return cache[1], cache[2]


class my_Bottle2neck(_CacheInvalidating):
    """Linear(in -> w*s) -> split into s groups of width w; group i (i < s-1) is passed through its own
    Linear+BN+ReLU after adding the previous group's output (hierarchical residual); the last group is
    passed through unchanged; concat -> Linear(w*s -> planes) + BN, residual (optionally projected), ReLU."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, baseWidth=26, scale=4, stype='normal'):
        super().__init__()
        width = int(math.floor(planes * (baseWidth / 64.0)))
        self.conv1 = nn.Linear(inplanes, width * scale, bias=False)
        self.bn1 = nn.BatchNorm1d(width * scale)
        self.nums = 1 if scale == 1 else scale - 1
        if stype == 'stage':
            self.pool = nn.AvgPool1d(kernel_size=3, stride=stride, padding=1)
        self.convs = nn.ModuleList([nn.Linear(width, width, bias=False) for _ in range(self.nums)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(width) for _ in range(self.nums)])
        self.conv3 = nn.Linear(width * scale, planes, bias=False)
        self.bn3 = nn.BatchNorm1d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stype = stype
        self.scale = scale
        self.width = width

    def _chain_pack(self):
        """Folded weights / shifts of convs[i] / bns[i], packed once for kpreg_chain_forward (cached until a
        parameter or running statistic changes)."""
        from . import ops
        folded = [_folded(self.convs[i], self.bns[i]) for i in range(self.nums)]
        key = tuple(self.bns[i]._kpreg_folded[0] for i in range(self.nums))
        cache = getattr(self, "_kpreg_chain", None)
        if cache is None or cache[0] != key:
            cache = (key, ops.ChainPack(torch.stack([f[0] for f in folded]), torch.stack([f[1] for f in folded])))
            self._kpreg_chain = cache
        return cache[1]

    def _pair_weight(self, i, wt):
        """[W_i | 0 | W_i] of chain layer i for ops.linear_pair_forward, cached with the folded weight it is built from."""
        from . import ops
        key = self.bns[i]._kpreg_folded[0]
        cache = getattr(self.bns[i], "_kpreg_pair", None)
        if cache is None or cache[0] != key:
            cache = (key, ops.pair_weight(wt))
            self.bns[i]._kpreg_pair = cache
        return cache[1]

    def _front_pack(self):
        """Folded conv1 / bn1 and convs[i] / bns[i] as the operand boxes of kpreg_front_forward (cached like _chain_pack)."""
        from . import ops
        w1, b1 = _folded(self.conv1, self.bn1)
        folded = [_folded(self.convs[i], self.bns[i]) for i in range(self.nums)]
        key = (self.bn1._kpreg_folded[0],) + tuple(self.bns[i]._kpreg_folded[0] for i in range(self.nums))
        cache = getattr(self, "_kpreg_front", None)
        if cache is None or cache[0] != key:
            cache = (key, ops.FrontPack(w1, b1, torch.stack([f[0] for f in folded]), torch.stack([f[1] for f in folded])))
            self._kpreg_front = cache
        return cache[1]

    def _fused_forward(self, x, shortcut=None):
        """Inference on CUDA: every Linear + eval-BatchNorm (+ ReLU) is one tensor-core GEMM with a fused
        epilogue; the chained layers run in one register-resident kernel where the width allows it (otherwise
        each emits `its output + the next group`, the next layer's input); the residual projection is folded
        into the last GEMM by concatenating its input and weights along K; ``shortcut`` (the enclosing block's
        identity shortcut) makes that GEMM return leaky_relu(unit output + shortcut, 0.1)."""
        from . import kpconv_blocks as kb
        from . import ops
        w, n_groups = self.width, self.scale
        gemm = kb.DEFAULT_GEMM
        wt, sh = _folded(self.conv1, self.bn1)
        fuse_res = self.downsample is not None and x.shape[1] % 4 == 0
        k_cat = w * n_groups
        front = (kb.FRONT_KERNEL and gemm == 1 and self.nums == n_groups - 1 and x.dtype == torch.float32 and x.stride(1) == 1
                 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and ops.front_supported(w, n_groups, x.shape[1]))
        chain = front or (kb.CHAIN_KERNEL and self.nums == n_groups - 1 and ops.chain_supported(w, self.nums))
        # The front / chain kernels write the copy of x behind the concatenation themselves.  The layer-by-layer path (wide units)
        # leaves x where it is: conv3 + the residual projection are then ONE GEMM whose reduction runs over z and over x.
        res_pair = (fuse_res and not chain and gemm == 1 and kb.PAIR_CONV3 and k_cat % 4 == 0
                    and ops.linear_pair_supported(x, x, self.conv3.weight.shape[0]))
        z = torch.empty((x.shape[0], k_cat + (x.shape[1] if fuse_res and not res_pair else 0)), dtype=x.dtype, device=x.device)
        if front:
            # conv1 and all chained layers in one tcgen05 kernel: conv1's output never leaves the SM
            ops.front_forward(x, self._front_pack(), z, copy_x=fuse_res)
            t = None
        else:
            # Layer by layer: conv1 writes its 8 groups straight into z and every chained layer overwrites "its" group with
            # its output — the pass-through group is then already in place.  Safe where ONE output tile of k_gemm_tc covers
            # all w columns (kpreg_linear_tile_cols: 128 at w = 112, 256 at w = 224 with the fp16 split): a tile's rows are
            # stored after its whole reduction has been read, and no other tile reads those rows.  Otherwise: separate t and
            # one copy of the last group.
            in_place = (not chain and gemm == 1 and kb.PAIR_CONV3 and w % 4 == 0 and self.nums == n_groups - 1 and self.nums > 1
                        and w <= min(ops.linear_tile_cols(w, w), ops.linear_tile_cols((w + 31) // 32 * 32 + w, w)))
            t = ops.linear_forward(x, wt, None, sh, act="relu", gemm=gemm, out=z[:, :k_cat] if in_place else None)  # [N, w * scale]
        if front:
            pass
        elif chain:
            # all chained layers in one kernel: a warp keeps its rows of the running activation in registers,
            # t is read once, z written once (the pass-through group and the copy of x included)
            ops.chain_forward(t, self._chain_pack(), z, x if fuse_res else None)
        else:
            # layer by layer.  Layer i >= 1 multiplies (out_{i-1} + t_i) by its matrix: on the tensor-core path that is ONE
            # GEMM whose reduction runs over out_{i-1} (read from z, where layer i-1 wrote it) and then over t_i, with the
            # matrix stacked twice along K — no `out + next group` side output, no addend read in the epilogue.
            pair = gemm == 1 and w % 4 == 0 and self.nums > 1 and ops.linear_pair_supported(z[:, :w], t[:, w:2 * w], w)
            scratch = None if pair else [torch.empty((t.shape[0], w), dtype=t.dtype, device=t.device) for _ in range(2)]
            inp = t[:, :w]
            for i in range(self.nums):
                wt, sh = _folded(self.convs[i], self.bns[i])
                nxt = i + 1 < self.nums
                if pair and i >= 1:
                    ops.linear_pair_forward(z[:, (i - 1) * w:i * w], t[:, i * w:(i + 1) * w], self._pair_weight(i, wt), sh,
                                            act="relu", out=z[:, i * w:(i + 1) * w])
                    continue
                ops.linear_forward(inp, wt, None, sh, act="relu", out=z[:, i * w:(i + 1) * w],
                                   out2=scratch[i & 1] if (nxt and not pair) else None,
                                   addend=t[:, (i + 1) * w:(i + 2) * w] if (nxt and not pair) else None, gemm=gemm)
                if not pair:
                    inp = scratch[i & 1]
            if t.data_ptr() != z.data_ptr():
                z[:, self.nums * w:k_cat] = t[:, self.nums * w:]
        w3, b3 = _folded(self.conv3, self.bn3)
        if fuse_res:
            # relu(bn3(cat W3^T) + bn_d(x Wd^T)) = relu([cat | x] [W3' | Wd']^T + b3' + bd')
            if not chain and not res_pair:
                z[:, k_cat:] = x
            wd, bd = _folded(self.downsample[0], self.downsample[1])
            key = (self.bn3._kpreg_folded[0], self.downsample[1]._kpreg_folded[0], res_pair)
            cache = getattr(self, "_kpreg_joint", None)
            if cache is None or cache[0] != key:
                pad = (-k_cat) % 32 if res_pair else 0  # the pair GEMM's second operand starts at a multiple of 32 columns
                cache = (key, torch.cat([w3, w3.new_zeros((w3.shape[0], pad)), wd], 1).contiguous(), (b3 + bd).contiguous())
                self._kpreg_joint = cache
            if res_pair:
                return ops.linear_pair_forward(z, x, cache[1], cache[2], act="relu", post_residual=shortcut,
                                               post_act="leaky_relu" if shortcut is not None else None)
            return ops.linear_forward(z, cache[1], None, cache[2], act="relu", gemm=gemm, post_residual=shortcut,
                                      post_act="leaky_relu" if shortcut is not None else None)
        residual = x
        if self.downsample is not None:
            wd, bd = _folded(self.downsample[0], self.downsample[1])
            residual = ops.linear_forward(x, wd, None, bd, gemm=gemm)
        return ops.linear_forward(z[:, :k_cat], w3, None, b3, residual=residual, act="relu", gemm=gemm, post_residual=shortcut,
                                  post_act="leaky_relu" if shortcut is not None else None)

    def forward(self, x, shortcut=None):
        """``shortcut`` (optional, not in the reference): returns leaky_relu(unit(x) + shortcut, 0.1), the tail of
        ResnetBottleneckBlock.forward, fused into the last GEMM on the inference path."""
        if (not self.training and x.is_cuda and not torch.is_grad_enabled() and self.stype == 'normal'
                and self.scale > 1 and self.width % 4 == 0):
            from . import kpconv_blocks as kb
            if kb.FUSED_GLUE:
                return self._fused_forward(x, shortcut)
        from . import ops
        lin = ops.linear_train  # bias-free nn.Linear under autograd: tensor-core forward / dx / d_weight where the shape allows
        groups = torch.split(self.relu(self.bn1(lin(x, self.conv1))), self.width, 1)
        outs, carry = [], None
        for i in range(self.nums):
            carry = groups[i] if (i == 0 or self.stype == 'stage') else carry + groups[i]
            carry = self.relu(self.bns[i](lin(carry.contiguous(), self.convs[i])))
            outs.append(carry)
        if self.scale != 1:
            outs.append(groups[self.nums] if self.stype == 'normal' else self.pool(groups[self.nums]))
        out = self.bn3(lin(torch.cat(outs, 1), self.conv3))
        residual = x if self.downsample is None else self.downsample[1](lin(x, self.downsample[0]))
        out = self.relu(out + residual)
        return out if shortcut is None else torch.nn.functional.leaky_relu(out + shortcut, 0.1)


class my_res2Net(_CacheInvalidating):
    """One ``block`` mapping in_dim -> out_dim channels, with a Linear+BN projection on the residual."""

    def __init__(self, block, in_dim, out_dim, baseWidth=26, scale=4):
        super().__init__()
        self.inplanes = in_dim
        self.out_dim = out_dim
        self.baseWidth = baseWidth
        self.scale = scale
        downsample = None
        if self.inplanes != out_dim * block.expansion:
            downsample = nn.Sequential(nn.Linear(self.inplanes, self.out_dim, bias=False),
                                       nn.BatchNorm1d(self.out_dim))
        self.layer1 = nn.Sequential(block(self.inplanes, out_dim, 1, downsample=downsample, stype='normal',
                                          baseWidth=baseWidth, scale=scale))

    def forward(self, x, shortcut=None):
        return self.layer1(x) if shortcut is None else self.layer1[0](x, shortcut)
