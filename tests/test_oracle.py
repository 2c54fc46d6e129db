"""Pins the CPU oracle (oracle/) against the reference: the compiled reference C++ (oracle/_ref) and
the golden vectors produced by the reference's own Python (tests/golden/make_golden.py)."""
import os
import re

import numpy as np
import pytest
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import kpconv_config, synthetic
from conftest import ROOT, rel_err

LEVEL_KEYS = ("points", "neighbors", "pools", "upsamples", "stack_lengths")


def _have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libkpref.so"))


def assert_rows_equal_up_to_ties(got, want, ties, what=""):
    """Index tables must be bit-identical, except inside rows that hold two exactly equal d2: there the
    reference's order is whatever its unstable std::sort leaves — a 2-point voxel's
    barycentre is equidistant from both points, so such rows do occur in `pools` — and only the row's
    SET of indices is defined."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, what
    ties = np.asarray(ties, bool)
    assert np.array_equal(got[~ties], want[~ties]), what
    assert np.array_equal(np.sort(got[ties], 1), np.sort(want[ties], 1)), what


def tie_rows_from_reference(oracle, q, s, ql, sl, radius, chunk=65536):
    """Rows holding two exactly equal d2, from the compiled reference's own (untruncated, ascending-d2) rows: d2 is
    recomputed in fp32 in nanoflann's order ((dx*dx + dy*dy) + dz*dz, no FMA — numpy rounds every operation) and equal
    values are adjacent in a sorted row.  Linear in the table size; the port's detector is an O(Nq x Ns) sweep."""
    q, s = np.ascontiguousarray(q, np.float32), np.ascontiguousarray(s, np.float32)
    rows = oracle.batch_query(q, s, ql, sl, radius, impl="ref")
    n_s = s.shape[0]
    s_pad = np.concatenate([s, np.full((1, 3), np.float32(1e18))], 0)
    ties = np.zeros(q.shape[0], bool)
    for a in range(0, q.shape[0], chunk):
        idx = rows[a:a + chunk]
        d = q[a:a + chunk, None, :] - s_pad[idx]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        valid = idx < n_s
        ties[a:a + chunk] = ((d2[:, 1:] == d2[:, :-1]) & valid[:, 1:] & valid[:, :-1]).any(1)
    return ties


def tie_rows(oracle, q, s, ql, sl, radius):
    if oracle.have_ref() and float(len(q)) * float(len(s)) > 2e9:
        return tie_rows_from_reference(oracle, q, s, ql, sl, radius)  # full-size clouds: skip the port's quadratic sweep
    return oracle.batch_query(q, s, ql, sl, radius, impl="port", return_ties=True)[1]


def check_pyramid(oracle, meta, want, cfg):
    """meta / want: dicts of per-level lists.  Points and lengths bit-exact; tables exact up to ties."""
    n_levels = len(want["points"])
    r = cfg.first_subsampling_dl * cfg.conv_radius
    for lvl in range(n_levels):
        assert np.array_equal(np.asarray(meta["points"][lvl]), want["points"][lvl]), ("points", lvl)
        assert np.array_equal(np.asarray(meta["stack_lengths"][lvl]).astype(np.int64),
                              np.asarray(want["stack_lengths"][lvl]).astype(np.int64)), ("stack_lengths", lvl)
    for lvl in range(n_levels):
        p, l = want["points"][lvl], want["stack_lengths"][lvl]
        assert_rows_equal_up_to_ties(meta["neighbors"][lvl], want["neighbors"][lvl], tie_rows(oracle, p, p, l, l, r),
                                     ("neighbors", lvl))
        if lvl + 1 < n_levels:
            p2, l2 = want["points"][lvl + 1], want["stack_lengths"][lvl + 1]
            assert_rows_equal_up_to_ties(meta["pools"][lvl], want["pools"][lvl], tie_rows(oracle, p2, p, l2, l, r),
                                         ("pools", lvl))
            assert_rows_equal_up_to_ties(meta["upsamples"][lvl], want["upsamples"][lvl],
                                         tie_rows(oracle, p, p2, l, l2, 2 * r), ("upsamples", lvl))
        else:
            assert np.asarray(meta["pools"][lvl]).shape == (0, 1) and np.asarray(meta["upsamples"][lvl]).shape == (0, 1)
        r *= 2


def golden_pyramid_3dmatch(g):
    want = {key: [g.get(f"{key}_{lvl}") for lvl in range(4)] for key in LEVEL_KEYS}
    want["points"][0] = np.concatenate([g["src"], g["tgt"]], 0)
    return want


def _levels(g, prefix, key):
    out, lvl = [], 0
    while f"{prefix}{key}_{lvl}" in g:
        out.append(g[f"{prefix}{key}_{lvl}"])
        lvl += 1
    return out


def test_port_pyramid_matches_golden_modelnet(oracle, golden_modelnet):
    g = golden_modelnet
    cfg = kpconv_config("modelnet", first_feats_dim=64)
    meta = oracle.preprocess([g["mn_src"], g["mn_tgt"]], cfg, impl="port")
    want = {key: _levels(g, "mn_", key) for key in LEVEL_KEYS}
    assert all(len(want[k]) == len(meta[k]) == 2 for k in LEVEL_KEYS)
    for key in LEVEL_KEYS:
        for a, b in zip(meta[key], want[key]):
            assert a.dtype == b.dtype, key
    check_pyramid(oracle, meta, want, cfg)


def test_port_pyramid_matches_golden_3dmatch(oracle, golden_3dmatch):
    g = golden_3dmatch
    cfg = kpconv_config("3dmatch")
    meta = oracle.preprocess([g["src"], g["tgt"]], cfg, impl="port")
    assert len(meta["points"]) == 4
    check_pyramid(oracle, meta, golden_pyramid_3dmatch(g), cfg)


@pytest.mark.parametrize("seed,n,dl", [(0, 1, 0.1), (1, 13, 0.05), (2, 14, 0.05), (3, 500, 0.07), (4, 6000, 0.04),
                                       (5, 6000, 5.0), (6, 30000, 0.03)])
def test_port_subsample_matches_reference_cpp(oracle, seed, n, dl):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(seed)
    lens = np.array([n, max(1, n // 2), n], np.int32)
    pts = rng.uniform(-1, 1, size=(int(lens.sum()), 3)).astype(np.float32)
    a, la = oracle.subsample_batch(pts, lens, dl, impl="port")
    b, lb = oracle.subsample_batch(pts, lens, dl, impl="ref")
    assert np.array_equal(la, lb) and np.array_equal(a, b)
    # max_p keeps the first max_p of each cloud
    a2, la2 = oracle.subsample_batch(pts, lens, dl, max_p=7, impl="port")
    b2, lb2 = oracle.subsample_batch(pts, lens, dl, max_p=7, impl="ref")
    assert np.array_equal(la2, lb2) and np.array_equal(a2, b2)


@pytest.mark.parametrize("seed,nq,ns,r", [(0, 300, 300, 0.2), (1, 50, 900, 0.35), (2, 900, 40, 0.5), (3, 2000, 2000, 0.08)])
def test_port_neighbors_match_reference_cpp(oracle, seed, nq, ns, r):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(seed)
    ql = np.array([nq, nq // 3 + 1], np.int32)
    sl = np.array([ns, ns // 2 + 1], np.int32)
    q = rng.uniform(-1, 1, size=(int(ql.sum()), 3)).astype(np.float32)
    s = rng.uniform(-1, 1, size=(int(sl.sum()), 3)).astype(np.float32)
    a, ties = oracle.batch_query(q, s, ql, sl, r, impl="port", return_ties=True)
    b = oracle.batch_query(q, s, ql, sl, r, impl="ref")
    assert a.shape == b.shape
    assert np.array_equal(a[~ties], b[~ties])  # generic position: identical rows
    assert np.array_equal(np.sort(a, 1), np.sort(b, 1))  # tied rows: same sets


def test_neighbor_ties_are_sets_only(oracle):
    """On exact-tie data (a lattice) the reference's order is unspecified (unstable sort over KD-tree
    visit order): only the neighbour SETS are comparable, and the port flags those rows."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    g = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.1
    lens = np.array([g.shape[0]], np.int32)
    a, ties = oracle.batch_query(g, g, lens, lens, 0.15, impl="port", return_ties=True)
    b = oracle.batch_query(g, g, lens, lens, 0.15, impl="ref")
    assert ties.any()
    assert np.array_equal(np.sort(a, 1), np.sort(b, 1))


@pytest.mark.parametrize("infl", ["linear", "gaussian", "constant"])
@pytest.mark.parametrize("agg", ["sum", "closest"])
def test_oracle_kpconv_matches_reference_python(oracle, golden_modelnet, infl, agg):
    g = golden_modelnet
    out = oracle.kpconv_forward(g["mn_points_1"], g["mn_points_0"], g["mn_pools_0"], g["op_x"],
                                g[f"op_{infl}_{agg}_w"], g[f"op_{infl}_{agg}_kp"], 0.12, infl, agg)
    assert rel_err(out.numpy(), g[f"op_{infl}_{agg}_out"]) < 1e-5


def test_oracle_max_pool_matches_reference_python(oracle, golden_modelnet):
    g = golden_modelnet
    assert np.array_equal(oracle.max_pool(g["op_x"], g["mn_pools_0"]).numpy(), g["op_maxpool"])


def test_oracle_encoder_matches_reference_python(oracle, golden_modelnet):
    g = golden_modelnet
    cfg = kpconv_config("modelnet", first_feats_dim=64)
    sd = {k[len("mn_sd::"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("mn_sd::")}
    batch = {key: _levels(g, "mn_", key) for key in LEVEL_KEYS}
    x0 = np.ones((g["mn_points_0"].shape[0], 1), np.float32)
    y, skips = oracle.encoder_forward(sd, cfg, x0, batch)
    assert y.shape == g["mn_enc_out"].shape
    assert rel_err(y.numpy(), g["mn_enc_out"]) < 1e-4
    assert len(skips) == 2 and rel_err(skips[1].numpy(), g["mn_enc_skip_1"]) < 1e-4


def test_oracle_kabsch_matches_reference_python(oracle, golden_modelnet):
    g = golden_modelnet
    for key, w in (("kb_T_weighted", g["kb_w"]), ("kb_T_unweighted", None), ("kb_T_zero", np.zeros_like(g["kb_w"]))):
        t = oracle.compute_rigid_transform(g["kb_a"], g["kb_b"], w)
        err = oracle.pose_error(t, torch.from_numpy(g[key]))
        assert float(err["rot_deg"].max()) < 1e-3 and float(err["trans"].max()) < 1e-5, key
    t = oracle.fast_compute_rigid_transform(g["kb_a"], g["kb_b"], g["kb_w"], 0.85)
    err = oracle.pose_error(t, torch.from_numpy(g["kb_T_fast"]))
    assert float(err["rot_deg"].max()) < 1e-3 and float(err["trans"].max()) < 1e-5
    # and the recovered pose is the generating pose up to the injected noise
    err = oracle.se3_compare(torch.from_numpy(g["kb_T_weighted"]), torch.from_numpy(g["kb_pose"]).expand(6, 3, 4))
    assert float(err["rot_deg"].max()) < 0.2


def test_cuda_table_sizes_match_libstdcxx_header():
    """The bucket-count table hard-coded in the CUDA kernel is the one derived from libstdc++."""
    hdr = open(os.path.join(ROOT, "oracle", "prime_growth.h")).read()
    sizes = [int(x) for x in re.findall(r"^\s*(\d+)ull,", hdr, flags=re.M)]
    pkg = os.path.dirname(kpreg_b200.__file__)
    cu = open(os.path.join(pkg, "csrc", "subsample.cu")).read()
    block = cu[cu.index("c_table_sizes[27]"):]
    block = block[:block.index("};")]
    cuda_sizes = [int(x) for x in re.findall(r"(\d+)u", block)]
    assert cuda_sizes == sizes[:27]


@pytest.mark.skipif(not _have_ref(), reason="needs the compiled reference (oracle/_ref)")
def test_tie_detector_from_reference_rows_matches_port():
    """The linear-time tie detector used for full-size clouds marks exactly the rows the port's exhaustive sweep marks."""
    import kp_oracle
    from kpreg_b200 import synthetic
    src, tgt, _ = synthetic.modelnet_pair(seed=5)
    pts = np.concatenate([src, tgt], 0)
    lens = np.array([len(src), len(tgt)], np.int32)
    sub, sub_l = kp_oracle.subsample_batch(pts, lens, sampleDl=0.06, impl="port")
    for q, s, ql, sl, r in ((pts, pts, lens, lens, 0.0825), (sub, pts, sub_l, lens, 0.0825), (pts, sub, lens, sub_l, 0.165)):
        want = kp_oracle.batch_query(q, s, ql, sl, r, impl="port", return_ties=True)[1]
        got = tie_rows_from_reference(kp_oracle, q, s, ql, sl, r)
        assert np.array_equal(got, want)
    # a lattice: ties everywhere
    g = np.stack(np.meshgrid(*[np.arange(6, dtype=np.float32) * 0.05] * 3, indexing="ij"), -1).reshape(-1, 3)
    gl = np.array([len(g)], np.int32)
    want = kp_oracle.batch_query(g, g, gl, gl, 0.12, impl="port", return_ties=True)[1]
    assert want.any() and np.array_equal(tie_rows_from_reference(kp_oracle, g, g, gl, gl, 0.12), want)


def r2_state_dict(g, prefix, enc):
    """Seeded weights (tests/seeded_weights.py) + the fixture's kernel points for an encoder of the fixture's shape."""
    import torch as _t
    from seeded_weights import seeded_state_dict
    sd = seeded_state_dict(enc.state_dict())
    for k in sd:
        if k.endswith("kernel_points"):
            sd[k] = _t.from_numpy(g[f"{prefix}_kp::{k}"])
    return sd


def r2_case(g, prefix):
    """(cfg, d_bottle, [src, tgt]) of an encoder_r2.npz case."""
    if prefix == "tdm":
        cfg = kpconv_config("3dmatch")
        src, tgt, _ = synthetic.threedmatch_pair(seed=2, n_raw=9000)
        return cfg, cfg.d_embed, [src, tgt]
    cfg = kpconv_config("modelnet", first_feats_dim=128)
    src, tgt, _ = synthetic.modelnet_pair(seed=1)
    return cfg, 256, [src, tgt]


@pytest.mark.parametrize("prefix", ["tdm", "mn128"])
def test_oracle_encoder_matches_reference_r2_fixture(oracle, golden_encoder_r2, prefix):
    """The port's encoder against the REFERENCE KPFEncoder on the shipped 3DMatch configuration (res2net widths
    28 / 56 / 112 / 224) and on the ModelNet architecture at width 28 — the widths this repo's fused path serves."""
    from kpreg_b200.kpconv import KPFEncoder
    g = golden_encoder_r2
    cfg, d_bottle, clouds = r2_case(g, prefix)
    np.random.seed(0)
    sd = r2_state_dict(g, prefix, KPFEncoder(cfg, d_bottle))
    meta = oracle.preprocess(clouds, cfg, impl="ref" if oracle.have_ref() else "port")
    x0 = np.ones((meta["points"][0].shape[0], 1), np.float32)
    y, skips = oracle.encoder_forward(sd, cfg, x0, meta)
    assert y.shape == g[f"{prefix}_enc_out"].shape
    err = rel_err(y.numpy(), g[f"{prefix}_enc_out"])
    stride = int(g[f"{prefix}_row_stride"])
    errs = [rel_err(s.numpy()[::stride], g[f"{prefix}_skip_{i}_rows"]) for i, s in enumerate(skips)]
    print(f"oracle port vs reference KPFEncoder ({prefix}): out {err:.2e}, skips {['%.1e' % e for e in errs]}")
    assert err < 1e-4 and max(errs) < 1e-4


def test_oracle_compute_overlaps_matches_reference_fixture(oracle, golden_3dmatch, golden_overlaps_r2):
    g, ov = golden_3dmatch, golden_overlaps_r2
    want = golden_pyramid_3dmatch(g)
    pyr = oracle.compute_overlaps(ov["src_overlap"], ov["tgt_overlap"], want["pools"], [p.shape[0] for p in want["points"]])
    assert len(pyr) == 4
    for p in range(4):
        assert float(np.abs(pyr[p] - ov[f"pyr_{p}"]).max()) < 1e-6, p  # sums of <= 40 fp32 terms: order-of-summation slack only
