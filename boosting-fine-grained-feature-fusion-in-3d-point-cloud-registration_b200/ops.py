"""Device-tensor operators over the C ABI (``include/kpreg_b200.h``).

Everything here takes and returns CUDA tensors and enqueues work on the current CUDA stream
without synchronising the host.  Shapes that depend on the data (number of subsampled points,
neighbour row widths) are returned as small device tensors; the callers in ``cpp_wrappers.py`` and
``kpconv.py`` decide when to read them.
"""
from __future__ import annotations

import contextlib
import functools
from torch.utils.weak import WeakIdKeyDictionary
from typing import Optional, Tuple

import torch

from . import _lib


def _on_tensor_device(fn):
    """Run ``fn`` with the CUDA device of its first tensor argument current.  The C entry points launch on the current
    device, so tensors living on another GPU of the same process (one process driving several GPUs) need the switch;
    when the devices already agree — one process per GPU, the normal case — this is one integer comparison."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor):
                if a.is_cuda and a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapper

INFLUENCE = {"constant": 0, "linear": 1, "gaussian": 2}
AGGREGATION = {"sum": 0, "closest": 1}


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _lib.require_cuda(t, name)
    return t.detach().to(torch.float32).contiguous()


def _i32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _lib.require_cuda(t, name)
    return t.detach().to(torch.int32).contiguous()


def _idx(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    _lib.require_cuda(t, name)
    if t.dtype == torch.int64:
        return t.contiguous(), 1
    if t.dtype == torch.int32:
        return t.contiguous(), 0
    return t.to(torch.int64).contiguous(), 1


_NULL_CTX = contextlib.nullcontext()


def _device_of(t: torch.Tensor):
    """Context manager making ``t``'s CUDA device current (a no-op object when it already is)."""
    if isinstance(t, torch.Tensor) and t.is_cuda and t.device.index != torch.cuda.current_device():
        return torch.cuda.device(t.device)
    return _NULL_CTX


# ------------------------------------------------------------------------------------------------
# subsampling
# ------------------------------------------------------------------------------------------------

@_on_tensor_device
def subsample(points: torch.Tensor, lens: torch.Tensor, sample_dl: float, max_p: int = 0):
    """Voxel-grid barycentres of a stacked batch (kpreg_subsample_batch).

    Returns (out_pts [N,3] — only the first M rows are valid, counts int32 [B+2] =
    per-cloud counts, M, status).  Nothing is synchronised.
    """
    lib = _lib.load()
    points = _f32c(points, "points")
    lens = _i32c(lens, "lens")
    n, b = points.shape[0], lens.shape[0]
    dev = points.device
    out = torch.empty((max(n, 1), 3), dtype=torch.float32, device=dev)
    counts = torch.empty(b + 2, dtype=torch.int32, device=dev)
    nbytes = _lib.size_query("kpreg_subsample_workspace_bytes", n, b)
    ws = _lib.workspaces.get(nbytes, dev)
    rc = lib.kpreg_subsample_batch(points.data_ptr(), lens.data_ptr(), n, b, float(sample_dl), int(max_p),
                                   out.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(),
                                   _lib.stream_ptr(dev))
    _lib.check(rc, "kpreg_subsample_batch")
    return out[:n], counts


# ------------------------------------------------------------------------------------------------
# neighbours
# ------------------------------------------------------------------------------------------------

class CellGrid:
    """Supports binned into cells of edge ``cell`` (kpreg_grid_build); serves radius queries with
    radius <= cell for any query set of the same batch (kpreg_grid_query)."""

    def __init__(self, supports: torch.Tensor, s_lens: torch.Tensor, cell: float, want_order: bool = True):
        with _device_of(supports):
            self._build(supports, s_lens, cell, want_order)

    def _build(self, supports, s_lens, cell, want_order):
        lib = _lib.load()
        self.supports = _f32c(supports, "supports")
        self.s_lens = _i32c(s_lens, "s_batches")
        self.n = int(self.supports.shape[0])
        self.n_clouds = int(self.s_lens.shape[0])
        self.cell = float(cell)
        dev = self.supports.device
        nbytes = _lib.size_query("kpreg_grid_workspace_bytes", self.n, self.n_clouds)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        # cell-sorted permutation of the supports: a spatially coherent processing order for later kernels
        self.order = torch.empty(self.n, dtype=torch.int32, device=dev) if want_order and self.n > 0 else None
        rc = lib.kpreg_grid_build(self.supports.data_ptr(), self.s_lens.data_ptr(), self.n, self.n_clouds,
                                  self.cell, self.buf.data_ptr(), self.buf.numel(), _lib.ptr(self.order),
                                  _lib.stream_ptr(dev))
        _lib.check(rc, "kpreg_grid_build")

    def query(self, queries: torch.Tensor, q_lens: torch.Tensor, radius: float, width: int,
              stats: Optional[torch.Tensor] = None, want_counts: bool = False, idx64: bool = False,
              order: Optional[torch.Tensor] = None):
        """Rows of ascending-(d2, index) neighbours, truncated/padded to ``width`` columns.

        Returns (idx [Nq,width], counts [Nq] or None, stats int32 [2] = {max count, status})."""
        with _device_of(queries):
            return self._query(queries, q_lens, radius, width, stats, want_counts, idx64, order)

    def _query(self, queries, q_lens, radius, width, stats, want_counts, idx64, order):
        lib = _lib.load()
        queries = _f32c(queries, "queries")
        q_lens = _i32c(q_lens, "q_batches")
        if q_lens.shape[0] != self.n_clouds:
            raise RuntimeError("queries and supports must have the same number of clouds")
        dev = queries.device
        nq = int(queries.shape[0])
        out = torch.empty((nq, width), dtype=torch.int64 if idx64 else torch.int32, device=dev)
        counts = torch.empty(nq, dtype=torch.int32, device=dev) if want_counts else None
        if stats is None:
            stats = torch.zeros(2, dtype=torch.int32, device=dev)
        rc = lib.kpreg_grid_query(self.buf.data_ptr(), self.n, self.n_clouds, queries.data_ptr(), q_lens.data_ptr(),
                                  nq, float(radius), int(width), 1 if idx64 else 0, _order_ptr(order, nq),
                                  out.data_ptr(), _lib.ptr(counts), stats.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, "kpreg_grid_query")
        return out, counts, stats


def _order_ptr(order: Optional[torch.Tensor], n_rows: int):
    """Pointer of an optional int32 processing-order permutation of n_rows query rows."""
    if order is None:
        return None
    if not order.is_cuda or order.dtype != torch.int32 or order.numel() != n_rows or not order.is_contiguous():
        raise RuntimeError("order must be a contiguous int32 CUDA permutation of the query rows")
    return order.data_ptr()


@_on_tensor_device
def pack_rows(rows: torch.Tensor, out_width: int, idx64: bool) -> torch.Tensor:
    """[n, in_width] int32 -> [n, out_width] int32/int64 (kpreg_pack_rows)."""
    lib = _lib.load()
    _lib.require_cuda(rows, "rows")
    rows = rows.contiguous()
    n, in_w = rows.shape
    out = torch.empty((n, out_width), dtype=torch.int64 if idx64 else torch.int32, device=rows.device)
    rc = lib.kpreg_pack_rows(rows.data_ptr(), n, in_w, int(out_width), 1 if idx64 else 0, out.data_ptr(),
                             _lib.stream_ptr(rows.device))
    _lib.check(rc, "kpreg_pack_rows")
    return out


# ------------------------------------------------------------------------------------------------
# KPConv
# ------------------------------------------------------------------------------------------------

@_on_tensor_device
def kpconv_forward(q_pts, s_pts, idx, x, weights, kernel_points, kp_extent: float, influence: str = "linear",
                   aggregation: str = "sum", gemm: int = 0, order: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    if influence not in INFLUENCE:
        raise ValueError("Unknown influence function type (config.KP_influence)")
    if aggregation not in AGGREGATION:
        raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")
    row_pos = row_predicate_of(x)  # written by the segment norm that produced x (inference), or None
    q_pts, s_pts, x = _f32c(q_pts, "q_pts"), _f32c(s_pts, "s_pts"), _f32c(x, "x")
    weights_in = weights
    weights, kernel_points = _f32c(weights, "weights"), _f32c(kernel_points, "kernel_points")
    idx, idx64 = _idx(idx, "neighb_inds")
    n_q, n_s, h = q_pts.shape[0], s_pts.shape[0], idx.shape[1] if idx.dim() == 2 else 0
    k, c_in, c_out = weights.shape
    if x.shape[0] != n_s or x.shape[1] != c_in or idx.shape[0] != n_q or kernel_points.shape[0] != k:
        raise RuntimeError("KPConv: inconsistent shapes")
    if row_pos is not None and (row_pos[1] != x.data_ptr() or row_pos[0].shape[0] != n_s or row_pos[0].device != x.device):
        row_pos = None
    dev = q_pts.device
    out = torch.empty((n_q, c_out), dtype=torch.float32, device=dev)
    nbytes = _lib.size_query("kpreg_kpconv_workspace_bytes", n_q, n_s, k, c_in, c_out, 0)
    ws = _lib.workspaces.get(nbytes, dev)
    w_ptr = weights.data_ptr()
    if gemm == 1 and not torch.is_grad_enabled() and (k * c_in) % 4 == 0 and k * c_in >= 4 and c_out >= 8 and n_q > 0:
        w_ptr, gemm = cached_split(weights_in, True).buf.data_ptr(), 2  # inference: the weights were split once
    rc = lib.kpreg_kpconv_forward_rowpos(q_pts.data_ptr(), s_pts.data_ptr(), idx.data_ptr(), idx64, x.data_ptr(),
                                         w_ptr, kernel_points.data_ptr(), n_q, n_s, h, k, c_in, c_out,
                                         float(kp_extent), INFLUENCE[influence], AGGREGATION[aggregation], int(gemm),
                                         _order_ptr(order, n_q), None if row_pos is None else row_pos[0].data_ptr(),
                                         out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "kpreg_kpconv_forward")
    return out


def row_predicate_of(x):
    """(flags uint8 [N], data_ptr) attached by segment_norm(..., row_pos=True) to its output, if ``x`` is that very tensor
    and has not been written since (version counter); None otherwise."""
    tag = getattr(x, "_kpreg_row_pos", None)
    if tag is None or x.is_inference() or tag[2] != x._version:
        return None
    return tag[0], tag[1]


@_on_tensor_device
def kpconv_backward(q_pts, s_pts, idx, x, weights, kernel_points, grad_out, kp_extent: float,
                    influence: str = "linear", aggregation: str = "sum", order: Optional[torch.Tensor] = None):
    """Returns (d_x [n_s,c_in], d_weights [K,c_in,c_out])."""
    lib = _lib.load()
    q_pts, s_pts, x = _f32c(q_pts, "q_pts"), _f32c(s_pts, "s_pts"), _f32c(x, "x")
    weights, kernel_points = _f32c(weights, "weights"), _f32c(kernel_points, "kernel_points")
    grad_out = _f32c(grad_out, "grad_out")
    idx, idx64 = _idx(idx, "neighb_inds")
    n_q, n_s, h = q_pts.shape[0], s_pts.shape[0], idx.shape[1] if idx.dim() == 2 else 0
    k, c_in, c_out = weights.shape
    dev = q_pts.device
    d_x = torch.empty((n_s, c_in), dtype=torch.float32, device=dev)
    d_w = torch.empty((k, c_in, c_out), dtype=torch.float32, device=dev)
    nbytes = _lib.size_query("kpreg_kpconv_workspace_bytes", n_q, n_s, k, c_in, c_out, 1)
    ws = _lib.workspaces.get(nbytes, dev)
    rc = lib.kpreg_kpconv_backward(q_pts.data_ptr(), s_pts.data_ptr(), idx.data_ptr(), idx64, x.data_ptr(),
                                   weights.data_ptr(), kernel_points.data_ptr(), grad_out.data_ptr(), n_q, n_s, h, k,
                                   c_in, c_out, float(kp_extent), INFLUENCE[influence], AGGREGATION[aggregation],
                                   _order_ptr(order, n_q), d_x.data_ptr(), d_w.data_ptr(), ws.data_ptr(), ws.numel(),
                                   _lib.stream_ptr(dev))
    _lib.check(rc, "kpreg_kpconv_backward")
    return d_x, d_w


# ------------------------------------------------------------------------------------------------
# max pool
# ------------------------------------------------------------------------------------------------

@_on_tensor_device
def max_pool_forward(x, idx, want_argmax: bool = False, order: Optional[torch.Tensor] = None):
    lib = _lib.load()
    x = _f32c(x, "x")
    idx, idx64 = _idx(idx, "inds")
    n_s, c = x.shape
    n_q, h = idx.shape
    out = torch.empty((n_q, c), dtype=torch.float32, device=x.device)
    arg = torch.empty((n_q, c), dtype=torch.int32, device=x.device) if want_argmax else None
    rc = lib.kpreg_max_pool_forward(x.data_ptr(), idx.data_ptr(), idx64, n_q, n_s, h, c, _order_ptr(order, n_q),
                                    out.data_ptr(), _lib.ptr(arg), _lib.stream_ptr(x.device))
    _lib.check(rc, "kpreg_max_pool_forward")
    return out, arg


@_on_tensor_device
def max_pool_backward(grad_out, argmax, n_s: int):
    lib = _lib.load()
    grad_out = _f32c(grad_out, "grad_out")
    n_q, c = grad_out.shape
    d_x = torch.empty((n_s, c), dtype=torch.float32, device=grad_out.device)
    rc = lib.kpreg_max_pool_backward(grad_out.data_ptr(), argmax.data_ptr(), n_q, n_s, c, d_x.data_ptr(),
                                     _lib.stream_ptr(grad_out.device))
    _lib.check(rc, "kpreg_max_pool_backward")
    return d_x


# ------------------------------------------------------------------------------------------------
# Kabsch
# ------------------------------------------------------------------------------------------------

@_on_tensor_device
def kabsch(a: torch.Tensor, b: torch.Tensor, w: Optional[torch.Tensor], n_sets: int, pts_per_set: int,
           offsets: Optional[torch.Tensor] = None, threshold: float = -1.0, write_back: bool = False) -> torch.Tensor:
    """a, b [total,3] f32 (contiguous), w [total] f32 or None -> [n_sets,3,4] (kpreg_kabsch).

    ``w`` is modified in place when write_back is set (it must then be contiguous float32)."""
    lib = _lib.load()
    a, b = _f32c(a, "a"), _f32c(b, "b")
    if w is not None:
        _lib.require_cuda(w, "weights")
        if w.dtype != torch.float32 or not w.is_contiguous():
            if write_back:
                raise RuntimeError("in-place weight thresholding needs a contiguous float32 tensor")
            w = w.to(torch.float32).contiguous()
    out = torch.empty((n_sets, 3, 4), dtype=torch.float32, device=a.device)
    rc = lib.kpreg_kabsch(a.data_ptr(), b.data_ptr(), _lib.ptr(w), _lib.ptr(offsets), int(n_sets), int(pts_per_set),
                          float(threshold), 1 if write_back else 0, out.data_ptr(), _lib.stream_ptr(a.device))
    _lib.check(rc, "kpreg_kabsch")
    return out


# ------------------------------------------------------------------------------------------------
# encoder-block glue: Linear (+ folded BatchNorm + activation) and per-cloud instance norm
# ------------------------------------------------------------------------------------------------

ACT = {None: 0, "none": 0, "relu": 1, "leaky_relu": 2}


class SplitWeights:
    """hi/lo TF32 operand pair of a weight matrix, produced once by kpreg_split_weights (inference)."""

    def __init__(self, weight: torch.Tensor, transpose: bool):
        lib = _lib.load()
        w = _f32c(weight, "weight")
        if transpose:  # KPConv weights [K, c_in, c_out] -> [K*c_in, c_out]
            w = w.reshape(-1, w.shape[-1])
            self.k, self.n = int(w.shape[0]), int(w.shape[1])
        else:
            self.n, self.k = int(w.shape[0]), int(w.shape[1])
        nbytes = _lib.size_query("kpreg_linear_workspace_bytes", self.k, self.n)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
        rc = lib.kpreg_split_weights(w.data_ptr(), self.k, self.n, 1 if transpose else 0, self.buf.data_ptr(), nbytes,
                                     _lib.stream_ptr(w.device))
        _lib.check(rc, "kpreg_split_weights")


# Inference caches (split weights here; folded Linear+BatchNorm weights, chain packs and joint conv3/downsample matrices
# in res2net.py) are validated by tensor identity + Tensor._version + this epoch.  In-place updates through the autograd
# API (optimizer steps, load_state_dict, copy_) bump _version; writes through ``.data`` (``p.data.copy_()``, hand-written
# EMA updates, ``bn.running_var.data.fill_()``) do NOT — call ``invalidate_caches()`` after such writes.  The encoder's
# modules call it themselves from train() / eval(), load_state_dict() and .to() / .cuda() / .float().
_CACHE_EPOCH = [0]
_SPLIT_CACHE = WeakIdKeyDictionary()  # weight tensor -> {transpose: (version, epoch, SplitWeights)}


def cache_epoch() -> int:
    return _CACHE_EPOCH[0]


def invalidate_caches() -> None:
    """Drop every cached derived weight (split TF32 operands, folded Linear+BatchNorm, chain packs).  Needed only after
    parameter / running-statistic writes that bypass autograd's version counter (``.data`` writes)."""
    _CACHE_EPOCH[0] += 1
    _SPLIT_CACHE.clear()


class CacheInvalidatingModule(torch.nn.Module):
    """nn.Module whose mode switches, state-dict loads and device / dtype moves invalidate the inference caches."""

    def train(self, mode: bool = True):
        invalidate_caches()
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        invalidate_caches()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        invalidate_caches()
        return super()._load_from_state_dict(*args, **kwargs)


def cached_split(weight: torch.Tensor, transpose: bool) -> SplitWeights:
    """Split-weight cache for inference, keyed weakly by the weight tensor itself (a dead tensor's entry — and its
    device buffer — disappears with it) and validated by its version counter and the cache epoch."""
    per_tensor = _SPLIT_CACHE.get(weight)
    if per_tensor is None:
        per_tensor = {}
        _SPLIT_CACHE[weight] = per_tensor
    hit = per_tensor.get(bool(transpose))
    if hit is None or hit[0] != weight._version or hit[1] != _CACHE_EPOCH[0]:
        hit = (weight._version, _CACHE_EPOCH[0], SplitWeights(weight, transpose))
        per_tensor[bool(transpose)] = hit
    return hit[2]


def _rows(t: torch.Tensor, name: str):
    """(tensor, row pitch) of a 2-D fp32 CUDA tensor whose rows are contiguous (column slices are fine)."""
    _lib.require_cuda(t, name)
    if t.dtype != torch.float32 or t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        t = t.to(torch.float32).contiguous()
    return t, int(t.stride(0)) if t.shape[0] > 1 else max(int(t.stride(0)), int(t.shape[1]))


@_on_tensor_device
def linear_forward(x, weight, col_scale=None, col_shift=None, residual=None, act=None, slope: float = 0.1, out=None,
                   out2=None, addend=None, gemm: int = 1, post_residual=None, post_act=None):
    """act((x @ weight.T) * col_scale + col_shift + residual) -> out [M,N] (kpreg_linear_forward).
    ``out`` may be a column slice of a wider buffer; ``out2`` (optional) receives out + addend;
    ``post_residual`` (optional) is added after ``act`` and followed by ``post_act``."""
    lib = _lib.load()
    x, ldx = _rows(x, "x")
    m, k = x.shape
    presplit = None
    # the TMA path's addressing rule (kpreg_gemm_supported): K >= 4, N >= 8, 16-byte aligned base and row pitch
    if (gemm == 1 and not torch.is_grad_enabled() and m > 0 and k >= 4 and weight.shape[0] >= 8 and ldx % 4 == 0
            and x.data_ptr() % 16 == 0):
        presplit = cached_split(weight, False)  # inference: the weights were split once
    weight = _f32c(weight, "weight")
    n = weight.shape[0]
    if weight.shape[1] != k:
        raise RuntimeError("linear: inconsistent shapes")
    dev = x.device
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=dev)
    ldc = int(out.stride(0)) if m > 1 else max(int(out.stride(0)), n)
    res_ptr, ld_res = None, 0
    if residual is not None:
        residual, ld_res = _rows(residual, "residual")
        res_ptr = residual.data_ptr()
    o2_ptr, ld2, add_ptr, ld_add = None, 0, None, 0
    if out2 is not None:
        addend, ld_add = _rows(addend, "addend")
        o2_ptr, ld2, add_ptr = out2.data_ptr(), int(out2.stride(0)), addend.data_ptr()
    post_ptr, ld_post = None, 0
    if post_residual is not None:
        post_residual, ld_post = _rows(post_residual, "post_residual")
        post_ptr = post_residual.data_ptr()
    cs = None if col_scale is None else _f32c(col_scale, "col_scale")
    cb = None if col_shift is None else _f32c(col_shift, "col_shift")
    nbytes = _lib.size_query("kpreg_linear_workspace_bytes", k, n)
    ws = _lib.workspaces.get(nbytes, dev)
    w_ptr = presplit.buf.data_ptr() if presplit is not None else weight.data_ptr()
    rc = lib.kpreg_linear_forward(x.data_ptr(), ldx, w_ptr, m, k, n, _lib.ptr(cs), _lib.ptr(cb), res_ptr, ld_res,
                                  ACT[act], float(slope), out.data_ptr(), ldc, o2_ptr, ld2, add_ptr, ld_add, post_ptr, ld_post,
                                  ACT[post_act], 2 if presplit is not None else int(gemm), ws.data_ptr(), ws.numel(),
                                  _lib.stream_ptr(dev))
    _lib.check(rc, "kpreg_linear_forward")
    return out


def pair_weight(weight: torch.Tensor) -> torch.Tensor:
    """[W | zero columns up to a multiple of 32 | W] for ``linear_pair_forward`` with the same matrix on both inputs."""
    n, k = weight.shape
    pad = (-k) % 32
    return torch.cat([weight, weight.new_zeros((n, pad)), weight], 1).contiguous()


def linear_tile_cols(k: int, n: int) -> int:
    """Output columns per CTA tile of the tensor-core GEMM for a [*, k] x [k, n] product (kpreg_linear_tile_cols)."""
    return int(_lib.load().kpreg_linear_tile_cols(int(k), int(n)))


def linear_pair_supported(x1: torch.Tensor, x2: torch.Tensor, n: int) -> bool:
    """The TMA addressing rule of kpreg_linear_pair_forward: fp32 row slices, 16-byte aligned bases and row pitches."""
    ok = True
    for t in (x1, x2):
        ok = ok and (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 4 == 0
                     and t.data_ptr() % 16 == 0)
    return bool(ok and x1.shape[0] == x2.shape[0] and x1.shape[0] > 0 and x1.shape[1] >= 4 and n >= 8)


@_on_tensor_device
def linear_pair_forward(x1, x2, weight_cat, col_shift=None, act=None, slope: float = 0.1, out=None, col_scale=None,
                        post_residual=None, post_act=None):
    """post_act(act(([x1 | x2] @ weight_cat.T) * col_scale + col_shift) + post_residual) with the two inputs read where they
    lie (kpreg_linear_pair_forward); weight_cat = [W1 | 0 .. | W2] with W1 padded to a multiple of 32 columns."""
    lib = _lib.load()
    _lib.require_cuda(x1, "x1")
    m, k1 = x1.shape
    k2 = x2.shape[1]
    n = weight_cat.shape[0]
    if weight_cat.shape[1] != (k1 + 31) // 32 * 32 + k2 or not linear_pair_supported(x1, x2, n):
        raise RuntimeError("linear_pair: inconsistent shapes or unaligned inputs")
    split = cached_split(weight_cat, False)
    dev = x1.device
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=dev)
    ldc = int(out.stride(0)) if m > 1 else max(int(out.stride(0)), n)
    cs = None if col_scale is None else _f32c(col_scale, "col_scale")
    cb = None if col_shift is None else _f32c(col_shift, "col_shift")
    pr, ld_post = None, 0
    if post_residual is not None:
        pr, ld_post = _rows(post_residual, "post_residual")
        if pr.shape[0] != m or pr.shape[1] != n:
            raise RuntimeError("linear_pair: post_residual must be [M, N]")
    rc = lib.kpreg_linear_pair_forward(x1.data_ptr(), int(x1.stride(0)) if m > 1 else k1, k1, x2.data_ptr(),
                                       int(x2.stride(0)) if m > 1 else k2, k2, split.buf.data_ptr(), m, n, _lib.ptr(cs), _lib.ptr(cb),
                                       ACT[act], float(slope), _lib.ptr(pr), int(ld_post), ACT[post_act], out.data_ptr(), ldc,
                                       _lib.stream_ptr(dev))
    _lib.check(rc, "kpreg_linear_pair_forward")
    return out


@_on_tensor_device
def linear_backward(x, grad_out, weight, need_dx: bool = True, need_dw: bool = True):
    """(dx [M,K] or None, d_weight [N,K] or None) of y = x weight^T on the tensor cores (kpreg_linear_backward).
    Returns None when a shape is outside the TMA paths: the caller then uses torch."""
    lib = _lib.load()
    x, ldx = _rows(x, "x")
    grad_out, ldg = _rows(grad_out, "grad_out")
    weight = _f32c(weight, "weight")
    m, k = x.shape
    n = weight.shape[0]
    if (m == 0 or k < 8 or n < 8 or ldx % 4 or ldg % 4 or x.data_ptr() % 16 or grad_out.data_ptr() % 16 or k % 4 or n % 4):
        return None
    dev = x.device
    dx = torch.empty((m, k), dtype=torch.float32, device=dev) if need_dx else None
    dw = torch.empty((n, k), dtype=torch.float32, device=dev) if need_dw else None
    nbytes = _lib.size_query("kpreg_linear_backward_workspace_bytes", m, k, n)
    ws = _lib.workspaces.get(nbytes, dev)
    rc = lib.kpreg_linear_backward(x.data_ptr(), ldx, grad_out.data_ptr(), ldg, weight.data_ptr(), m, k, n, _lib.ptr(dx), k,
                                   _lib.ptr(dw), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    if rc == 1:
        return None
    _lib.check(rc, "kpreg_linear_backward")
    return dx, dw


class LinearFn(torch.autograd.Function):
    """y = x weight^T (no bias) with forward, dx and d_weight on the tcgen05 3xTF32 GEMMs — the training-mode stand-in
    for the nn.Linear layers of the encoder blocks (fp32-grade results, unlike TF32-allowed cuBLAS)."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.save_for_backward(x, weight)
        return linear_forward(x, weight, gemm=1)

    @staticmethod
    def backward(ctx, grad):
        x, weight = ctx.saved_tensors
        need_dx, need_dw = ctx.needs_input_grad
        res = linear_backward(x, grad, weight, need_dx, need_dw)
        if res is None:
            g = grad.to(torch.float32)
            return (g @ weight if need_dx else None), (g.t() @ x if need_dw else None)
        return res


def linear_train(x: torch.Tensor, linear: torch.nn.Linear) -> torch.Tensor:
    """``linear(x)`` for a bias-free nn.Linear under autograd: tensor-core kernels when the shape allows, torch otherwise."""
    if (torch.is_grad_enabled() and (x.requires_grad or linear.weight.requires_grad)
            and linear.bias is None and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.shape[0] > 0
            and x.shape[1] % 4 == 0 and x.shape[1] >= 8 and linear.out_features % 4 == 0 and linear.out_features >= 8):
        return LinearFn.apply(x, linear.weight)
    return linear(x)


@_on_tensor_device
def segment_norm(x, lens, residual=None, act=None, slope: float = 0.1, eps: float = 1e-5, out=None, row_pos: bool = False):
    """Per-cloud, per-channel (x - mean) * rstd (+ residual) (+ activation) (kpreg_segment_norm_forward).
    ``row_pos``: the output feeds a KPConv — where the kernel can (32 / 64 channels) it also writes that KPConv's row
    predicate (feature sum > 0) and tags the output with it, which saves kpconv_forward its own pass over the features."""
    lib = _lib.load()
    x, ldx = _rows(x, "x")
    lens = _i32c(lens, "stack_lengths")
    n, c = x.shape
    dev = x.device
    own_out = out is None  # only a tensor allocated here is tagged: nothing else can write it behind the version counter
    if out is None:
        out = torch.empty((n, c), dtype=torch.float32, device=dev)
    res_ptr, ld_res = None, 0
    if residual is not None:
        residual, ld_res = _rows(residual, "residual")
        res_ptr = residual.data_ptr()
    nbytes = _lib.size_query("kpreg_segment_norm_workspace_bytes", int(lens.shape[0]), c)
    ws = _lib.workspaces.get(nbytes, dev)
    flags = None
    # (tensors created under torch.inference_mode() have no version counter: not tagged, KPConv runs its own row pass)
    if row_pos and own_out and n > 0 and not out.is_inference() and lib.kpreg_segment_norm_rowpos_supported(c):
        flags = torch.empty((n,), dtype=torch.uint8, device=dev)
    rc = lib.kpreg_segment_norm_forward_rowpos(x.data_ptr(), ldx, lens.data_ptr(), int(lens.shape[0]), n, c, float(eps), res_ptr,
                                               ld_res, ACT[act], float(slope), out.data_ptr(), int(out.stride(0)) if n > 1 else c,
                                               None if flags is None else flags.data_ptr(), ws.data_ptr(), ws.numel(),
                                               _lib.stream_ptr(dev))
    _lib.check(rc, "kpreg_segment_norm_forward")
    if flags is not None:
        out._kpreg_row_pos = (flags, out.data_ptr(), out._version)
    return out


@_on_tensor_device
def segment_norm_backward(x, grad_out, lens, eps: float = 1e-5):
    """d/dx of the plain per-cloud norm (kpreg_segment_norm_backward)."""
    lib = _lib.load()
    x, ldx = _rows(x, "x")
    grad_out, ldg = _rows(grad_out, "grad_out")
    lens = _i32c(lens, "stack_lengths")
    n, c = x.shape
    dx = torch.empty((n, c), dtype=torch.float32, device=x.device)
    nbytes = _lib.size_query("kpreg_segment_norm_workspace_bytes", int(lens.shape[0]), c)
    ws = _lib.workspaces.get(nbytes, x.device)
    rc = lib.kpreg_segment_norm_backward(x.data_ptr(), ldx, grad_out.data_ptr(), ldg, lens.data_ptr(), int(lens.shape[0]), n, c,
                                         float(eps), dx.data_ptr(), c, ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device))
    _lib.check(rc, "kpreg_segment_norm_backward")
    return dx


class ChainPack:
    """The folded weights / shifts of a res2net chain arranged once in the fragment order kpreg_chain_forward reads."""

    def __init__(self, weights: torch.Tensor, shifts: torch.Tensor):
        lib = _lib.load()
        weights, shifts = _f32c(weights, "weights"), _f32c(shifts, "shifts")
        self.n_layers, self.width = int(weights.shape[0]), int(weights.shape[1])
        if weights.shape[2] != self.width or tuple(shifts.shape) != (self.n_layers, self.width):
            raise RuntimeError("chain: weights [L, w, w] and shifts [L, w] expected")
        nbytes = _lib.size_query("kpreg_chain_pack_bytes", self.width, self.n_layers)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=weights.device)
        rc = lib.kpreg_chain_pack(weights.data_ptr(), shifts.data_ptr(), self.width, self.n_layers, self.buf.data_ptr(), nbytes,
                                  _lib.stream_ptr(weights.device))
        _lib.check(rc, "kpreg_chain_pack")


def chain_supported(width: int, n_layers: int) -> bool:
    return bool(_lib.load().kpreg_chain_supported(int(width), int(n_layers)))


@_on_tensor_device
def chain_forward(t: torch.Tensor, pack: ChainPack, z: torch.Tensor, x_copy: Optional[torch.Tensor] = None) -> torch.Tensor:
    """res2net's chained layers over conv1's output t [M, (L+1) w] into z [M, >= (L+1) w (+ c_x)] (kpreg_chain_forward)."""
    lib = _lib.load()
    t, ld_t = _rows(t, "t")
    _lib.require_cuda(z, "z")
    if z.dtype != torch.float32 or z.dim() != 2 or z.stride(1) != 1:
        raise RuntimeError("chain: z must be a float32 matrix with contiguous rows")
    m = t.shape[0]
    x_ptr, ld_x, c_x = None, 0, 0
    if x_copy is not None:
        x_copy, ld_x = _rows(x_copy, "x_copy")
        x_ptr, c_x = x_copy.data_ptr(), int(x_copy.shape[1])
    rc = lib.kpreg_chain_forward(t.data_ptr(), ld_t, pack.buf.data_ptr(), pack.width, pack.n_layers, m, z.data_ptr(),
                                 int(z.stride(0)) if m > 1 else int(z.shape[1]), x_ptr, ld_x, c_x, _lib.stream_ptr(t.device))
    _lib.check(rc, "kpreg_chain_forward")
    return z


class FrontPack:
    """conv1 and chain weights of a res2net unit (BatchNorm folded in) as the operand boxes kpreg_front_forward copies."""

    def __init__(self, w1: torch.Tensor, b1: torch.Tensor, wc: torch.Tensor, bc: torch.Tensor):
        lib = _lib.load()
        w1, b1, wc, bc = _f32c(w1, "w1"), _f32c(b1, "b1"), _f32c(wc, "wc"), _f32c(bc, "bc")
        self.width, self.n_groups, self.c_in = int(wc.shape[1]), int(wc.shape[0]) + 1, int(w1.shape[1])
        if (tuple(w1.shape) != (self.n_groups * self.width, self.c_in) or tuple(b1.shape) != (self.n_groups * self.width,)
                or tuple(wc.shape) != (self.n_groups - 1, self.width, self.width) or tuple(bc.shape) != (self.n_groups - 1, self.width)):
            raise RuntimeError("front: w1 [G w, c_in], b1 [G w], wc [G-1, w, w], bc [G-1, w] expected")
        nbytes = _lib.size_query("kpreg_front_pack_bytes", self.width, self.n_groups, self.c_in)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=w1.device)
        rc = lib.kpreg_front_pack(w1.data_ptr(), b1.data_ptr(), wc.data_ptr(), bc.data_ptr(), self.width, self.n_groups, self.c_in,
                                  self.buf.data_ptr(), nbytes, _lib.stream_ptr(w1.device))
        _lib.check(rc, "kpreg_front_pack")


def front_supported(width: int, n_groups: int, c_in: int) -> bool:
    return bool(_lib.load().kpreg_front_supported(int(width), int(n_groups), int(c_in)))


@_on_tensor_device
def front_forward(x: torch.Tensor, pack: FrontPack, z: torch.Tensor, copy_x: bool = False) -> torch.Tensor:
    """conv1 + the chained layers of a res2net unit over x [M, c_in] into z [M, >= G w (+ c_in)] (kpreg_front_forward)."""
    lib = _lib.load()
    x, ld_x = _rows(x, "x")
    _lib.require_cuda(z, "z")
    if z.dtype != torch.float32 or z.dim() != 2 or z.stride(1) != 1 or x.shape[1] != pack.c_in or z.shape[0] != x.shape[0]:
        raise RuntimeError("front: x [M, c_in] and a float32 z [M, >= G w] with contiguous rows expected")
    m = x.shape[0]
    rc = lib.kpreg_front_forward(x.data_ptr(), ld_x, pack.c_in, pack.buf.data_ptr(), pack.width, pack.n_groups, m, z.data_ptr(),
                                 int(z.stride(0)) if m > 1 else int(z.shape[1]), 1 if copy_x else 0, _lib.stream_ptr(x.device))
    _lib.check(rc, "kpreg_front_forward")
    return z


# ------------------------------------------------------------------------------------------------
# the steps on either side of the path: overlap pyramid, coarse-level packing, point shuffling
# ------------------------------------------------------------------------------------------------

@_on_tensor_device
def overlap_pool(level: torch.Tensor, pools: torch.Tensor) -> torch.Tensor:
    """One level of compute_overlaps: clamp(mean of level over the valid entries of each pooling row, 0, 1)."""
    lib = _lib.load()
    level = _f32c(level, "overlap level")
    idx, idx64 = _idx(pools, "pools")
    n_q, h = idx.shape
    out = torch.empty(n_q, dtype=torch.float32, device=level.device)
    rc = lib.kpreg_overlap_pool(level.data_ptr(), idx.data_ptr(), idx64, n_q, int(level.shape[0]), h, out.data_ptr(),
                                _lib.stream_ptr(level.device))
    _lib.check(rc, "kpreg_overlap_pool")
    return out


@_on_tensor_device
def sine_embed(xyz: torch.Tensor, d_model: int, num_feats: int, scale: float, dim_t: torch.Tensor) -> torch.Tensor:
    """[..., n_dim] points -> [..., d_model] sine / cosine position code (kpreg_sine_embed)."""
    lib = _lib.load()
    _lib.require_cuda(xyz, "xyz")
    flat = xyz.detach().to(torch.float32).reshape(-1, xyz.shape[-1]).contiguous()
    out = torch.empty((flat.shape[0], d_model), dtype=torch.float32, device=flat.device)
    rc = lib.kpreg_sine_embed(flat.data_ptr(), int(flat.shape[0]), int(flat.shape[1]), int(d_model), int(num_feats), float(scale),
                              _f32c(dim_t, "dim_t").data_ptr(), out.data_ptr(), _lib.stream_ptr(flat.device))
    _lib.check(rc, "kpreg_sine_embed")
    return out.reshape(*xyz.shape[:-1], d_model)


@_on_tensor_device
def pack_coarse(feats: Optional[torch.Tensor], xyz: Optional[torch.Tensor], lens: torch.Tensor, max_len, d_model: int,
                num_feats: int = 0, scale: float = 1.0, dim_t: Optional[torch.Tensor] = None):
    """Padded (features, position embedding, padding masks) of the two halves of a stacked coarse level
    (kpreg_pack_coarse).  Returns (src_feats, tgt_feats, src_pe, tgt_pe, src_mask, tgt_mask); the feature / embedding
    entries are None when ``feats`` / ``xyz`` is None."""
    lib = _lib.load()
    lens = _i32c(lens, "stack_lengths")
    n_pairs = int(lens.shape[0]) // 2
    dev = lens.device
    ns_max, nt_max = int(max_len[0]), int(max_len[1])
    ld_f = 0
    if feats is not None:
        feats, ld_f = _rows(feats, "feats")
        if feats.shape[1] != d_model:
            raise RuntimeError("pack_coarse: feats must have d_model columns")
    if xyz is not None:
        xyz = _f32c(xyz, "xyz")

    def buf(n_max, on):
        return torch.empty((n_max, n_pairs, d_model), dtype=torch.float32, device=dev) if on else None

    src_f, tgt_f = buf(ns_max, feats is not None), buf(nt_max, feats is not None)
    src_pe, tgt_pe = buf(ns_max, xyz is not None), buf(nt_max, xyz is not None)
    src_m = torch.empty((n_pairs, ns_max), dtype=torch.bool, device=dev)
    tgt_m = torch.empty((n_pairs, nt_max), dtype=torch.bool, device=dev)
    nbytes = _lib.size_query("kpreg_pack_coarse_workspace_bytes", n_pairs)
    ws = _lib.workspaces.get(nbytes, dev)
    rc = lib.kpreg_pack_coarse(_lib.ptr(feats), ld_f, _lib.ptr(xyz), lens.data_ptr(), n_pairs, int(d_model), int(num_feats),
                               float(scale), _lib.ptr(None if dim_t is None else _f32c(dim_t, "dim_t")), ns_max, nt_max,
                               _lib.ptr(src_f), _lib.ptr(tgt_f), _lib.ptr(src_pe), _lib.ptr(tgt_pe), src_m.data_ptr(),
                               tgt_m.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "kpreg_pack_coarse")
    return src_f, tgt_f, src_pe, tgt_pe, src_m, tgt_m


@_on_tensor_device
def shuffle_gather(pts: torch.Tensor, mask: Optional[torch.Tensor], perm: torch.Tensor, want_reverse: bool = False):
    """(pts[perm], mask[perm] or None, reverse index int64 [n_in] or None, status int32 [1]) — kpreg_shuffle_gather."""
    lib = _lib.load()
    pts = _f32c(pts, "points")
    _lib.require_cuda(perm, "perm")
    perm = perm.to(torch.int64).contiguous()
    n_in, n_out = int(pts.shape[0]), int(perm.shape[0])
    dev = pts.device
    out = torch.empty((n_out, 3), dtype=torch.float32, device=dev)
    m_in = m_out = None
    if mask is not None:
        _lib.require_cuda(mask, "mask")
        m_in = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous()
        m_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
    rev = torch.empty(n_in, dtype=torch.int64, device=dev) if want_reverse else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.kpreg_shuffle_gather(pts.data_ptr(), _lib.ptr(m_in), perm.data_ptr(), n_out, n_in, out.data_ptr(), _lib.ptr(m_out),
                                  _lib.ptr(rev), status.data_ptr(), _lib.stream_ptr(dev))
    _lib.check(rc, "kpreg_shuffle_gather")
    if m_out is not None and mask.dtype == torch.bool:
        m_out = m_out.view(torch.bool)
    return out, m_out, rev, status


@_on_tensor_device
def remap_pairs(corr: torch.Tensor, rev_src: torch.Tensor, rev_tgt: torch.Tensor):
    """Correspondences [2, P] through two reverse indices -> (remapped [2, P] int64, keep [P] bool)."""
    lib = _lib.load()
    _lib.require_cuda(corr, "correspondences")
    corr = corr.to(torch.int64).contiguous()
    n = int(corr.shape[1])
    out = torch.empty_like(corr)
    keep = torch.empty(n, dtype=torch.bool, device=corr.device)
    rc = lib.kpreg_remap_pairs(corr.data_ptr(), n, rev_src.data_ptr(), int(rev_src.shape[0]), rev_tgt.data_ptr(),
                               int(rev_tgt.shape[0]), out.data_ptr(), keep.data_ptr(), _lib.stream_ptr(corr.device))
    _lib.check(rc, "kpreg_remap_pairs")
    return out, keep
