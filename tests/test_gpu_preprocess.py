"""-m gpu: subsampling, neighbour search and the whole pyramid on CUDA vs the CPU oracle.
Index tables and subsampled coordinates must be BIT-EXACT (ties: see assert_rows_equal_up_to_ties)."""
import numpy as np
import pytest
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import kpconv_config, synthetic
from kpreg_b200.cpp_wrappers import cpp_neighbors, cpp_subsampling
from kpreg_b200.kpconv import Preprocessor, batch_grid_subsampling_kpconv, batch_neighbors_kpconv
from gpu_util import (LEVEL_KEYS, _levels, assert_rows_equal_up_to_ties, check_pyramid, cuda, golden_pyramid_3dmatch,
                      meta_to_numpy)

pytestmark = pytest.mark.gpu


def _impl(oracle):
    return "ref" if oracle.have_ref() else "port"


@pytest.mark.parametrize("seed,lens,dl", [
    (0, [1], 0.1), (1, [13, 14, 15], 0.05), (2, [29, 30, 1, 2], 0.02), (3, [500, 257, 258], 0.07),
    (4, [6000, 3000], 0.04), (5, [6000, 100], 5.0), (6, [30000, 20000, 10000], 0.03), (7, [5] * 40, 0.3),
    (8, [120000], 0.021), (9, [100000, 5, 70000, 1], 0.03), (10, [40000, 300000], 0.012)])
def test_subsample_bit_exact(oracle, seed, lens, dl):
    rng = np.random.default_rng(seed)
    lens = np.array(lens, np.int32)
    pts = (rng.uniform(-1, 1, size=(int(lens.sum()), 3)) * rng.uniform(0.5, 3.0)).astype(np.float32)
    want, want_len = oracle.subsample_batch(pts, lens, dl, impl=_impl(oracle))
    got, got_len = cpp_subsampling.subsample_batch(pts, lens, sampleDl=dl)
    assert got.dtype == np.float32 and got_len.dtype == np.int32
    assert np.array_equal(got_len, want_len)
    assert np.array_equal(got, want)  # order and fp32 barycentres
    # device tensors in -> device tensors out, same values
    g2, l2 = cpp_subsampling.subsample_batch(cuda(pts), cuda(lens), sampleDl=dl)
    assert g2.is_cuda and np.array_equal(g2.cpu().numpy(), want) and np.array_equal(l2.cpu().numpy(), want_len)
    # max_p keeps the first max_p points of each cloud
    w3, wl3 = oracle.subsample_batch(pts, lens, dl, max_p=9, impl=_impl(oracle))
    g3, l3 = cpp_subsampling.subsample_batch(pts, lens, sampleDl=dl, max_p=9)
    assert np.array_equal(l3, wl3) and np.array_equal(g3, w3)


def test_subsample_voxel_property_full_size(oracle):
    """Size-independent property at the benchmark's size: every barycentre lies in the voxel it stands
    for, voxels are unique per cloud, and counts add up."""
    src, tgt, _ = synthetic.threedmatch_pair(seed=11)
    pts = np.concatenate([src, tgt])
    lens = np.array([len(src), len(tgt)], np.int32)
    dl = 0.05
    got, got_len = cpp_subsampling.subsample_batch(pts, lens, sampleDl=dl)
    assert got_len.sum() == got.shape[0]
    start_in, start_out = 0, 0
    for n_in, n_out in zip(lens, got_len):
        cloud, sub = pts[start_in:start_in + n_in], got[start_out:start_out + n_out]
        org = np.floor(cloud.min(0) * np.float32(1 / np.float32(dl))) * np.float32(dl)
        vox_in = np.unique(np.floor((cloud - org) / np.float32(dl)).astype(np.int64), axis=0)
        vox_out = np.unique(np.floor((sub - org) / np.float32(dl)).astype(np.int64), axis=0)
        assert len(vox_in) == n_out
        # barycentres sit inside the voxel they stand for (a handful may round onto a face)
        inside = len(set(map(tuple, vox_in.tolist())) & set(map(tuple, vox_out.tolist())))
        assert inside >= 0.999 * n_out
        start_in += n_in
        start_out += n_out


@pytest.mark.parametrize("seed,ql,sl,r", [
    (0, [300, 101], [300, 151], 0.2), (1, [50, 17], [900, 451], 0.35), (2, [900, 301], [40, 21], 0.5),
    (3, [2000, 667], [2000, 1001], 0.08), (4, [1], [1], 0.1), (5, [64, 64, 64, 64], [200, 3, 1, 777], 0.6)])
def test_batch_query_bit_exact(oracle, seed, ql, sl, r):
    rng = np.random.default_rng(seed)
    ql, sl = np.array(ql, np.int32), np.array(sl, np.int32)
    q = rng.uniform(-1, 1, size=(int(ql.sum()), 3)).astype(np.float32)
    s = rng.uniform(-1, 1, size=(int(sl.sum()), 3)).astype(np.float32)
    if len(q) == 1:
        s[0] = q[0] + np.float32(0.01)  # a lone query must have a neighbour (an empty result raises, as in the reference)
    want, ties = oracle.batch_query(q, s, ql, sl, r, impl="port", return_ties=True)
    got = cpp_neighbors.batch_query(q, s, ql, sl, radius=r)
    assert got.dtype == np.int32
    assert_rows_equal_up_to_ties(got, want, ties)
    if oracle.have_ref():
        assert_rows_equal_up_to_ties(got, oracle.batch_query(q, s, ql, sl, r, impl="ref"), ties)
    # truncation as in batch_neighbors_kpconv
    lim = max(1, want.shape[1] // 2)
    got_t = batch_neighbors_kpconv(cuda(q), cuda(s), cuda(ql), cuda(sl), r, lim)
    assert got_t.is_cuda
    assert_rows_equal_up_to_ties(got_t.cpu().numpy(), want[:, :lim], ties)


def test_batch_query_many_hits_uses_exact_overflow_path(oracle):
    """More neighbours per row than the shared-memory hit buffer (256): the recount path must give the
    same rows."""
    rng = np.random.default_rng(3)
    lens = np.array([1500, 900], np.int32)
    p = rng.uniform(-1, 1, size=(int(lens.sum()), 3)).astype(np.float32)
    want, ties = oracle.batch_query(p, p, lens, lens, 1.2, impl="port", return_ties=True)
    assert want.shape[1] > 300
    got = cpp_neighbors.batch_query(p, p, lens, lens, radius=1.2)
    assert_rows_equal_up_to_ties(got, want, ties)


def test_batch_query_lattice_ties_as_sets(oracle):
    g = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.1
    lens = np.array([g.shape[0]], np.int32)
    want, ties = oracle.batch_query(g, g, lens, lens, 0.15, impl="port", return_ties=True)
    got = cpp_neighbors.batch_query(g, g, lens, lens, radius=0.15)
    assert np.array_equal(got, want)  # the CUDA path and the port share the (d2, index) tie rule
    assert ties.any()


def test_batch_query_properties_full_size():
    """Benchmark-size properties: rows ascending in distance, all within the radius, first neighbour of a
    self-query is the point itself, and the relation is symmetric."""
    src, tgt, _ = synthetic.threedmatch_pair(seed=12)
    pts = np.concatenate([src, tgt])
    lens = np.array([len(src), len(tgt)], np.int32)
    r = 0.0625
    nb = cpp_neighbors.batch_query(pts, pts, lens, lens, radius=r)
    n = pts.shape[0]
    assert np.array_equal(nb[:, 0], np.arange(n))
    valid = nb < n
    padded = np.concatenate([pts, np.full((1, 3), 1e6, np.float32)])
    d = np.linalg.norm(padded[nb] - pts[:, None, :], axis=2)
    assert (d[valid] < r * (1 + 1e-6)).all()
    dd = np.where(valid, d, np.float32(1e3) + np.arange(nb.shape[1], dtype=np.float32)[None, :])
    assert (np.diff(dd, axis=1) >= -1e-7).all()
    rows = np.repeat(np.arange(n), nb.shape[1])[valid.ravel()]
    cols = nb.ravel()[valid.ravel()]
    fwd = set(zip(rows.tolist(), cols.tolist()))
    assert all((c, r_) in fwd for r_, c in list(fwd)[:20000])


def test_empty_result_raises_like_reference():
    q = np.zeros((3, 3), np.float32)
    s = np.ones((3, 3), np.float32) * 10
    lens = np.array([3], np.int32)
    with pytest.raises(RuntimeError):
        cpp_neighbors.batch_query(q, s, lens, lens, radius=0.1)
    with pytest.raises(RuntimeError):
        cpp_neighbors.batch_query(np.zeros((3, 2), np.float32), s, lens, lens, radius=0.1)
    with pytest.raises(RuntimeError):
        cpp_neighbors.batch_query(q, s, lens, np.array([1, 2], np.int32), radius=0.1)


def test_wrappers_return_torch(oracle):
    src, tgt, _ = synthetic.modelnet_pair(seed=4)
    pts = torch.from_numpy(np.concatenate([src, tgt]))
    lens = torch.tensor([len(src), len(tgt)], dtype=torch.int32)
    sp, sl = batch_grid_subsampling_kpconv(pts, lens, sampleDl=0.06)
    want, want_len = oracle.subsample_batch(pts, lens, 0.06, impl=_impl(oracle))
    assert isinstance(sp, torch.Tensor) and np.array_equal(sp.numpy(), want) and np.array_equal(sl.numpy(), want_len)


def test_pyramid_matches_golden_modelnet(oracle, golden_modelnet):
    g = golden_modelnet
    cfg = kpconv_config("modelnet", first_feats_dim=64)
    meta = Preprocessor(cfg)([cuda(g["mn_src"]), cuda(g["mn_tgt"])])
    for key in ("neighbors", "pools", "upsamples"):
        assert all(t.dtype == torch.int64 and t.is_cuda for t in meta[key])
    want = {key: _levels(g, "mn_", key) for key in LEVEL_KEYS}
    check_pyramid(oracle, meta_to_numpy(meta), want, cfg)
    # CPU tensors in -> CPU tensors out (the reference Preprocessor's contract)
    meta_cpu = Preprocessor(cfg)([torch.from_numpy(g["mn_src"]), torch.from_numpy(g["mn_tgt"])])
    assert all(not t.is_cuda for v in meta_cpu.values() for t in v)
    check_pyramid(oracle, meta_to_numpy(meta_cpu), want, cfg)


def test_pyramid_matches_golden_3dmatch(oracle, golden_3dmatch):
    g = golden_3dmatch
    cfg = kpconv_config("3dmatch")
    meta = Preprocessor(cfg)([cuda(g["src"]), cuda(g["tgt"])])
    check_pyramid(oracle, meta_to_numpy(meta), golden_pyramid_3dmatch(g), cfg)


@pytest.mark.parametrize("name,gen,n_pairs", [("3dmatch", synthetic.threedmatch_pair, 2), ("modelnet", synthetic.modelnet_pair, 3)])
def test_pyramid_matches_oracle_full_size(oracle, name, gen, n_pairs):
    """BASELINE config sizes, several pairs stacked (first all sources, then all targets)."""
    cfg = kpconv_config(name)
    pairs = [gen(seed=20 + i) for i in range(n_pairs)]
    clouds = [p[0] for p in pairs] + [p[1] for p in pairs]
    want = oracle.preprocess(clouds, cfg, impl=_impl(oracle))
    meta = Preprocessor(cfg, index_dtype=torch.int32)([cuda(c) for c in clouds])
    assert all(t.dtype == torch.int32 for t in meta["neighbors"])
    check_pyramid(oracle, meta_to_numpy(meta), want, cfg)


def test_pyramid_mcd_shape_sparse_grid(oracle):
    """LiDAR-like extent (80 m at r = 0.0625 m): the cell grid must stay sparse (hash table)."""
    cfg = kpconv_config("mcd")
    src, tgt, _ = synthetic.mcd_pair(seed=1, n=30000)
    want = oracle.preprocess([src, tgt], cfg, impl=_impl(oracle))
    meta = Preprocessor(cfg)([cuda(src), cuda(tgt)])
    check_pyramid(oracle, meta_to_numpy(meta), want, cfg)


def test_calibrate_neighbors_matches_reference_procedure(oracle):
    """calibrate_neighbors (reference finegrained_kpconv.py:707-739) on a small synthetic dataset vs the same
    histogram / percentile procedure over the CPU oracle's tables."""
    from kpreg_b200.kpconv import calibrate_neighbors
    cfg = kpconv_config("modelnet")
    data = [dict(zip(("src_xyz", "tgt_xyz"), synthetic.modelnet_pair(seed=50 + i)[:2])) for i in range(3)]
    got = calibrate_neighbors(data, cfg, keep_ratio=0.8, samples_threshold=10 ** 9)
    hist_n = int(np.ceil(4 / 3 * np.pi * (cfg.deform_radius + 1) ** 3))
    wide = kpconv_config("modelnet", neighborhood_limits=[hist_n] * 2)
    hists = np.zeros((cfg.num_layers, hist_n), np.int64)
    for item in data:
        meta = oracle.preprocess([item["src_xyz"], item["tgt_xyz"]], wide, impl=_impl(oracle))
        for lvl, t in enumerate(meta["neighbors"][:cfg.num_layers]):
            hists[lvl] += np.bincount((t < t.shape[0]).sum(1), minlength=hist_n)[:hist_n]
    cumsum = np.cumsum(hists.T, axis=0)
    want = np.sum(cumsum < (0.8 * cumsum[hist_n - 1, :]), axis=0)
    assert np.array_equal(got, want)


def test_pyramid_architecture_ending_on_a_strided_block(oracle):
    """The schedule's corner case (reference finegrained_kpconv.py:340-409): the last block is strided, so its level
    carries pool and upsample tables but the pooled points themselves are not appended."""
    cfg = kpconv_config("modelnet", architecture=["simple", "resnetb", "resnetb_strided"], num_layers=1)
    src, tgt, _ = synthetic.modelnet_pair(seed=8)
    want = oracle.preprocess([src, tgt], cfg, impl=_impl(oracle))
    meta = meta_to_numpy(Preprocessor(cfg)([cuda(src), cuda(tgt)]))
    assert len(want["points"]) == len(meta["points"]) == 1
    for key in ("points", "stack_lengths"):
        assert np.array_equal(np.asarray(meta[key][0]), np.asarray(want[key][0]))
    r = cfg.first_subsampling_dl * cfg.conv_radius
    p, l = want["points"][0], want["stack_lengths"][0]
    sub, sub_l = oracle.subsample_batch(p, l, 2 * r / cfg.conv_radius, impl=_impl(oracle))
    from gpu_util import assert_rows_equal_up_to_ties
    from test_oracle import tie_rows
    assert_rows_equal_up_to_ties(meta["neighbors"][0], want["neighbors"][0], tie_rows(oracle, p, p, l, l, r))
    assert_rows_equal_up_to_ties(meta["pools"][0], want["pools"][0], tie_rows(oracle, sub, p, sub_l, l, r))
    assert_rows_equal_up_to_ties(meta["upsamples"][0], want["upsamples"][0], tie_rows(oracle, p, sub, l, sub_l, 2 * r))


@pytest.mark.gpu
def test_staged_pyramid_and_overlapped_path_match_the_plain_path():
    """Preprocessor.stages (finest level first, the rest resumable under another CUDA stream) and RegistrationPath's
    overlapped launch order return exactly what the one-shot path returns: same tables, same features, same poses."""
    import torch
    from kpreg_b200 import kpconv_config, synthetic
    from kpreg_b200.kpconv import Preprocessor
    from kpreg_b200.pipeline import RegistrationPath
    cfg = kpconv_config("3dmatch")
    dev = torch.device("cuda")
    pairs = [synthetic.threedmatch_pair(seed=40 + i, n_raw=6000) for i in range(3)]
    src = [torch.from_numpy(p[0]).to(dev) for p in pairs]
    tgt = [torch.from_numpy(p[1]).to(dev) for p in pairs]
    poses = torch.from_numpy(np.stack([p[2] for p in pairs])).to(dev)
    pre = Preprocessor(cfg, index_dtype=torch.int32)
    whole = pre(src + tgt)
    staged = list(pre.stages(src + tgt, staged=True))
    assert len(staged) == 2 and set(staged[0]) == {"points", "neighbors", "stack_lengths", "orders"}
    assert torch.equal(staged[0]["neighbors"][0], whole["neighbors"][0]) and torch.equal(staged[0]["points"][0], whole["points"][0])
    for key in ("points", "neighbors", "pools", "upsamples", "stack_lengths"):
        assert all(torch.equal(a, b) for a, b in zip(staged[1][key], whole[key])), key
    torch.manual_seed(0)
    np.random.seed(0)
    plain = RegistrationPath(cfg, index_dtype=torch.int32, weights_threshold=0.85, overlap=False).eval().to(dev)
    both = RegistrationPath(cfg, index_dtype=torch.int32, weights_threshold=0.85, overlap=True).eval().to(dev)
    both.load_state_dict(plain.state_dict())
    a = plain(src, tgt, poses)
    for _ in range(3):  # repeated steps: tensors of the previous step die while both streams are busy
        b = both(src, tgt, poses)
        torch.cuda.synchronize()
        assert torch.equal(a["feats"], b["feats"]) and torch.equal(a["poses"], b["poses"])
