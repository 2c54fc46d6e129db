"""KPConv pyramid construction and encoder — host-side mirror of the reference's
``models/backbone_kpconv/finegrained_kpconv.py`` on top of the CUDA operators.

Drop-in surface (same names, arguments, result layout):

* ``batch_grid_subsampling_kpconv`` (reference :175-215) and ``batch_neighbors_kpconv`` (:248-263);
* ``Preprocessor(cfg)(pts) -> {'points','neighbors','pools','upsamples','stack_lengths'}`` (:296-419):
  index tensors int64 padded with the support level's row count, row width ``min(max_count, limit)``,
  empty ``(0,1)`` / ``(0,3)`` placeholders at the last level.  The results are those of the reference's
  deterministic CPU ``Preprocessor`` (not of its order-nondeterministic ``PreprocessorGPU``, :422-542),
  bit for bit, but computed on the device: the whole pyramid is enqueued with one small host read-back
  per subsampled level (its point count decides the next level's launch sizes) and one at the end (the
  row widths), instead of the reference's device->host->device round trip of every table;
* ``KPFEncoder(config, d_bottle, increase_channel_when_downsample=True)(x, batch) -> (x, skip_x)`` (:22-95).
"""
from __future__ import annotations

import logging
from typing import List

import numpy as np
import os

import torch
import torch.nn as nn

from . import ops
from .cpp_wrappers import cpp_neighbors, cpp_subsampling
from .kpconv_blocks import block_decider

_logger = logging.getLogger(__name__)


def batch_grid_subsampling_kpconv(points, batches_len, features=None, labels=None, sampleDl=0.1, max_p=0, verbose=0,
                                  random_grid_orient=True):
    """Grid subsampling of a stacked batch -> (s_points, s_len) tensors on the device of ``points``."""
    if features is not None or labels is not None:
        raise NotImplementedError("batch_grid_subsampling_kpconv: features / labels are not used by the registration path")
    s_points, s_len = cpp_subsampling.subsample_batch(points, batches_len, sampleDl=sampleDl, max_p=max_p,
                                                      verbose=verbose)
    if isinstance(s_points, np.ndarray):
        return torch.from_numpy(s_points), torch.from_numpy(s_len)
    return s_points, s_len


def batch_neighbors_kpconv(queries, supports, q_batches, s_batches, radius, max_neighbors):
    """Radius neighbours of a stacked batch, truncated to the ``max_neighbors`` closest when > 0."""
    neighbors = cpp_neighbors.batch_query(queries, supports, q_batches, s_batches, radius=radius)
    if isinstance(neighbors, np.ndarray):
        neighbors = torch.from_numpy(neighbors)
    return neighbors[:, :max_neighbors] if max_neighbors > 0 else neighbors


class _Table:
    """A neighbour table still in its over-allocated [Nq, limit] int32 form, awaiting its row width."""
    __slots__ = ("rows", "limit", "slot")

    def __init__(self, rows, limit, slot):
        self.rows, self.limit, self.slot = rows, limit, slot


class Preprocessor(nn.Module):
    """Computes the metadata used for KPConv (pyramid points, neighbour / pool / upsample tables)."""

    def __init__(self, cfg, index_dtype: torch.dtype = torch.int64):
        super().__init__()
        self.cfg = cfg
        self.index_dtype = index_dtype

    @torch.no_grad()
    def forward(self, pts: List[torch.Tensor]):
        out = None
        for out in self.stages(pts, staged=False):
            pass
        return out

    def stages(self, pts: List[torch.Tensor], staged: bool = True):
        """The pyramid as a generator.  ``staged``: first yields the finest level alone — ``{'points': [p0], 'neighbors': [n0],
        'stack_lengths': [l0], 'orders': [o0]}``, everything the encoder blocks in front of the first strided block read — and
        then the complete dict.  A caller can enqueue those blocks and resume the generator under another CUDA stream
        (``pipeline.RegistrationPath``): the radius queries of the coarser levels are instruction-bound and leave HBM idle,
        the finest level's Linear / norm layers are the opposite.  ``forward`` is ``stages(pts, staged=False)`` run to its end."""
        cfg = self.cfg
        limits = cfg.neighborhood_limits
        arch = list(cfg.architecture)
        in_device = pts[0].device
        dev = in_device if in_device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())

        lens_host = [int(p.shape[0]) for p in pts]
        points = torch.cat([p.to(torch.float32) for p in pts], dim=0)
        if not points.is_cuda:
            points = points.pin_memory().to(dev, non_blocking=True)
        lens = torch.tensor(lens_host, dtype=torch.int32).to(dev, non_blocking=True)
        n_clouds = len(lens_host)

        # every table's {max count, status} pair lands in one stats tensor, read once at the end
        n_tables_max = 3 * len(arch) + 3
        stats = torch.zeros((n_tables_max, 2), dtype=torch.int32, device=dev)
        next_slot = [0]

        def query(grid, q_pts, q_lens, radius, limit, order=None):
            slot = next_slot[0]
            next_slot[0] += 1
            if limit > 0:
                rows, _, _ = grid.query(q_pts, q_lens, radius, limit, stats=stats[slot], order=order)
                return _Table(rows, limit, slot)
            # no limit: the width itself is data dependent -> count first (one extra read-back)
            _, _, st = grid.query(q_pts, q_lens, radius, 0, stats=stats[slot], order=order)
            width = int(st[0].item())
            rows, _, _ = grid.query(q_pts, q_lens, radius, max(width, 1), stats=stats[slot], order=order)
            return _Table(rows, max(width, 1), slot)

        r_normal = cfg.first_subsampling_dl * cfg.conv_radius
        level_points, level_lens, level_orders = [], [], []
        conv_tabs, pool_tabs, up_tabs = [], [], []
        first_table = [None]  # staged: the finest level's conv table, finished ahead of the others
        layer_blocks, layer = [], 0
        grid = None           # cell grid of the current level (built one level ahead, see below)
        pending_up = None     # (fine points, fine lens, radius, limit, fine order): answered by the next level's grid

        for block_i, block in enumerate(arch):
            if 'global' in block or 'upsample' in block:
                break
            strided = 'pool' in block or 'strided' in block
            if not strided:
                layer_blocks.append(block)
                if block_i < len(arch) - 1 and 'upsample' not in arch[block_i + 1]:
                    continue

            deform_conv = any('deformable' in b for b in layer_blocks[:-1])
            r_conv = r_normal * cfg.deform_radius / cfg.conv_radius if deform_conv else r_normal
            r_pool = r_normal * cfg.deform_radius / cfg.conv_radius if 'deformable' in block else r_normal
            # one cell grid per level serves the conv table, the pool table (coarse queries) and the previous
            # level's upsample table (fine queries, radius 2*r_prev = r of this level); its cell-sorted permutation
            # doubles as the spatially coherent processing order of this level's rows (batch['orders'])
            cell = max(r_conv if layer_blocks else 0.0, r_pool if strided else 0.0,
                       pending_up[2] if pending_up is not None else 0.0)
            if grid is None or grid.cell < cell:
                grid = ops.CellGrid(points, lens, cell) if cell > 0 else None
            order = grid.order if grid is not None else None

            sub = counts = counts_host = counts_ready = None
            if strided:
                # The subsample goes first and its sizes start their way to the host at once (pinned buffer + event): the radius
                # queries enqueued behind it do not depend on them, so the host learns the next level's size while the GPU is
                # still busy and the one unavoidable read-back per level no longer drains the stream.
                dl = 2 * r_normal / cfg.conv_radius
                sub, counts = ops.subsample(points, lens, dl)
                if counts.is_cuda and os.environ.get("KPREG_SYNC_COUNTS", "")[:1] != "1":  # (=1: blocking read-back, A/B measurements)
                    counts_host = torch.empty(counts.shape, dtype=counts.dtype, pin_memory=True)
                    counts_host.copy_(counts, non_blocking=True)
                    counts_ready = torch.cuda.Event()
                    counts_ready.record(torch.cuda.current_stream(counts.device))

            if pending_up is not None:
                up_tabs.append(query(grid, pending_up[0], pending_up[1], pending_up[2], pending_up[3], pending_up[4]))
                pending_up = None
            conv_tabs.append(query(grid, points, lens, r_conv, limits[layer], order) if layer_blocks else None)

            level_points.append(points)
            level_lens.append(lens)
            level_orders.append(order)
            if staged and layer == 0 and conv_tabs[0] is not None and in_device.type == "cuda":
                # the finest level is complete: its table's width is read back now (one small extra read-back)
                st0 = stats[conv_tabs[0].slot].cpu()
                if int(st0[1]) != 0:
                    raise RuntimeError("Preprocessor: cell grid too large to index")
                first_table[0] = self._finish(conv_tabs[0], int(st0[0]), dev)
                yield {'points': [points], 'neighbors': [first_table[0]], 'stack_lengths': [lens], 'orders': [order]}
            if strided:
                if counts_ready is not None:
                    counts_ready.synchronize()
                    host = counts_host
                else:
                    host = counts.cpu()
                if int(host[-1]) != 0:
                    raise RuntimeError("Preprocessor: voxel grid too large to index")
                m = int(host[-2])
                if m < 1:
                    raise RuntimeError("Error")
                pool_p, pool_b = sub[:m], counts[:n_clouds]
                # the next level's grid is built now (cell = its conv radius = 2 r) so that its order can already
                # drive the pool query of the coarse points against this level's grid
                next_grid = ops.CellGrid(pool_p, pool_b, 2 * r_pool)
                pool_tabs.append(query(grid, pool_p, pool_b, r_pool, limits[layer], next_grid.order))
                pending_up = (points, lens, 2 * r_pool, limits[layer], order)
                points, lens, grid = pool_p, pool_b, next_grid
            else:
                pool_tabs.append(None)
                up_tabs.append(None)
                grid = None
            r_normal *= 2
            layer += 1
            layer_blocks = []

        if pending_up is not None:
            # the architecture ended on a strided block: its upsample table is answered by the grid already built over
            # the pooled points; like the reference, no further level is appended (finegrained_kpconv.py:395-409)
            up_tabs.append(query(grid, pending_up[0], pending_up[1], pending_up[2], pending_up[3], pending_up[4]))
            pending_up = None

        host_stats = stats.cpu()
        if int(host_stats[:, 1].max()) != 0:
            raise RuntimeError("Preprocessor: cell grid too large to index")

        def finish(tab):
            if tab is None:
                return torch.zeros((0, 1), dtype=torch.int64, device=dev)
            if first_table[0] is not None and tab is conv_tabs[0]:
                return first_table[0]
            return self._finish(tab, int(host_stats[tab.slot, 0]), dev)

        n_levels = len(level_points)
        up_tabs += [None] * (n_levels - len(up_tabs))
        data = {
            'points': level_points,
            'neighbors': [finish(t) for t in conv_tabs],
            'pools': [finish(t) for t in pool_tabs],
            'upsamples': [finish(t) for t in up_tabs],
            'stack_lengths': level_lens,
        }
        if in_device.type != "cuda":
            yield {k: [t.to(in_device) for t in v] for k, v in data.items()}
            return
        # extra key (not in the reference's dict): per level, the cell-sorted permutation of its rows — the blocks
        # hand it to the CUDA kernels as a processing order; consumers that do not know the key are unaffected
        data['orders'] = level_orders
        yield data

    def _finish(self, tab, max_count, dev):
        """An over-allocated [Nq, limit] table packed to its row width min(max_count, limit) and index dtype."""
        idx64 = self.index_dtype == torch.int64
        width = min(max_count, tab.limit)
        if tab.rows.shape[0] < 1 or width < 1:
            raise RuntimeError("Error")  # empty result: cpp_neighbors/wrapper.cpp:201-205
        if width == tab.limit and not idx64:
            return tab.rows
        return ops.pack_rows(tab.rows, width, idx64)


# The shipped model imports ``PreprocessorGPU`` (models/finegrained_regtr.py:32); this implementation is on the device
# already and returns the deterministic CPU Preprocessor's results, so the same class serves both names.
PreprocessorGPU = Preprocessor
batch_grid_subsampling_kpconv_gpu = batch_grid_subsampling_kpconv
batch_neighbors_kpconv_gpu = batch_neighbors_kpconv


def compute_overlaps(batch):
    """Ground-truth overlap ratio per point and pyramid level (reference :545-571): level 0 is the given
    per-point overlap mask; each further level averages the previous one over the valid entries of its pooling
    rows (kpreg_overlap_pool: one masked gather-mean kernel per level, on the device the tables live on)."""
    meta = batch['kpconv_meta']
    dev = meta['points'][0].device
    level = torch.cat([o.to(dev) for o in batch['src_overlap'] + batch['tgt_overlap']], dim=0).type(torch.float)
    pyramid = {'pyr_0': level}
    for p in range(1, len(meta['points'])):
        level = ops.overlap_pool(level, meta['pools'][p - 1])
        pyramid[f'pyr_{p}'] = level
    return pyramid


def calibrate_neighbors(dataset, config, collate_fn=None, keep_ratio=0.8, samples_threshold=2000):
    """Neighbourhood limits for a dataset (reference :707-739): histogram the per-point neighbour counts of every
    pyramid level over the dataset (until each level has more than ``samples_threshold`` samples) and keep, per level,
    the count below which ``keep_ratio`` of the points fall.  The pyramids are built on the device with the row limit
    lifted to the reference's bound ceil(4/3 pi (deform_radius + 1)^3).

    ``dataset[i]`` is a pair in collate_pair's format ({'src_xyz': [N,3], 'tgt_xyz': [M,3]}), a (src, tgt) tuple or
    a single [N,3] cloud; ``collate_fn`` (optional) maps an item to a list of [N,3] tensors instead."""
    from .config import AttrDict
    hist_n = int(np.ceil(4 / 3 * np.pi * (config.deform_radius + 1) ** 3))
    n_layers = int(config.num_layers)
    wide = AttrDict(dict(config))
    wide['neighborhood_limits'] = [hist_n] * max(n_layers, len(config.neighborhood_limits))
    pre = Preprocessor(wide, index_dtype=torch.int32)
    hists = torch.zeros((n_layers, hist_n), dtype=torch.int64)
    dev = torch.device('cuda', torch.cuda.current_device())
    for i in range(len(dataset)):
        item = dataset[i]
        if collate_fn is not None:
            clouds = collate_fn(item)
        elif isinstance(item, dict):
            clouds = [item['src_xyz'], item['tgt_xyz']]
        elif isinstance(item, (tuple, list)):
            clouds = list(item[:2])
        else:
            clouds = [item]
        clouds = [torch.as_tensor(c, dtype=torch.float32).to(dev) for c in clouds]
        meta = pre(clouds)
        for lvl, table in enumerate(meta['neighbors'][:n_layers]):
            counts = (table < table.shape[0]).sum(dim=1)
            hists[lvl] += torch.bincount(counts, minlength=hist_n)[:hist_n].cpu()
        if int(hists.sum(dim=1).min()) > samples_threshold:
            break
    cumsum = torch.cumsum(hists.t(), dim=0)
    return (cumsum < keep_ratio * cumsum[hist_n - 1, :].to(torch.float64)).sum(dim=0).numpy()


class KPFEncoder(torch.nn.Module):
    def __init__(self, config, d_bottle, increase_channel_when_downsample=True):
        super().__init__()
        self.logger = logging.getLogger(__name__)

        octave = 0
        r = config.first_subsampling_dl * config.conv_radius
        in_dim = config.in_feats_dim
        out_dim = config.first_feats_dim

        self.encoder_blocks = nn.ModuleList()
        self.encoder_skip_dims = []
        self.encoder_skips = []

        block = None
        for block_i, block in enumerate(config.architecture):
            if ('equivariant' in block) and (not out_dim % 3 == 0):
                raise ValueError('Equivariant block but features dimension is not a factor of 3')
            if any(tag in block for tag in ('pool', 'strided', 'upsample', 'global')):
                self.encoder_skips.append(block_i)
                self.encoder_skip_dims.append(in_dim)
            if 'upsample' in block:
                break
            self.encoder_blocks.append(block_decider(block, r, in_dim, out_dim, octave, config, flag=True))
            in_dim = out_dim // 2 if 'simple' in block else out_dim
            if 'pool' in block or 'strided' in block:
                octave += 1
                r *= 2
                if increase_channel_when_downsample:
                    out_dim *= 2

        if block is not None and 'upsample' not in block:
            # no decoder: the last block is a skip position too
            self.encoder_skips.append(block_i)
            self.encoder_skip_dims.append(in_dim)

    def forward(self, x, batch):
        return self.forward_blocks(x, batch, 0, len(self.encoder_blocks), [])

    def first_strided_block(self) -> int:
        """Index of the first block that reads anything beyond the finest level (its pool table, the next level's points)."""
        for i in self.encoder_skips:
            return i
        return len(self.encoder_blocks)

    def forward_blocks(self, x, batch, start: int, stop: int, skip_x):
        """Blocks [start, stop) of ``forward`` — the finest level's blocks can run while the rest of the pyramid is still
        being built (``Preprocessor.stages``).  ``skip_x`` collects the skip features across calls."""
        for block_i in range(start, stop):
            if block_i in self.encoder_skips:
                skip_x.append(x)
            x = self.encoder_blocks[block_i](x, batch)
        return x, skip_x
