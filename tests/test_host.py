"""CPU-side checks of the host mirror of the reference API (no kernels run)."""
import json
import os

import numpy as np
import pytest
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import kpconv_config, synthetic
from kpreg_b200.kernel_points import load_kernels
from kpreg_b200.kpconv import KPFEncoder
from kpreg_b200.kpconv_blocks import KPConv, ResnetBottleneckBlock, SimpleBlock, _segment_instance_norm, block_decider
from kpreg_b200.pipeline import shard_pairs
from conftest import GOLDEN, rel_err


@pytest.mark.parametrize("name", ["3dmatch", "modelnet"])
def test_encoder_state_dict_matches_reference_names_and_shapes(name):
    """Reference checkpoints must load: same parameter / buffer names and shapes."""
    want = json.load(open(os.path.join(GOLDEN, "encoder_state_dict_shapes.json")))[name]
    np.random.seed(0)
    cfg = kpconv_config(name)
    enc = KPFEncoder(cfg, cfg.d_embed)
    got = {k: list(v.shape) for k, v in enc.state_dict().items()}
    assert got == want
    if name == "3dmatch":
        assert enc.encoder_skips == [2, 5, 8, 10] and enc.encoder_skip_dims == [128, 256, 512, 1024]
        kinds = [type(b) for b in enc.encoder_blocks]
        assert kinds[0] is SimpleBlock and all(k is ResnetBottleneckBlock for k in kinds[1:])
        extents = [round(b.KPConv.KP_extent, 6) for b in enc.encoder_blocks]
        assert extents == [0.05, 0.05, 0.05, 0.1, 0.1, 0.1, 0.2, 0.2, 0.2, 0.4, 0.4]


def test_kpconv_constructor_and_errors():
    np.random.seed(0)
    conv = KPConv(15, 3, 8, 16, 0.05, 0.0625)
    assert conv.weights.shape == (15, 8, 16) and conv.kernel_points.shape == (15, 3)
    assert conv.weights.requires_grad and not conv.kernel_points.requires_grad
    assert float(conv.kernel_points[0].norm()) < 0.0625 * 0.05  # 'center': point 0 stays at the origin (+noise)
    assert float(conv.kernel_points.norm(dim=1).max()) < 0.0625
    with pytest.raises(ValueError):
        KPConv(15, 3, 8, 16, 0.05, 0.0625, KP_influence="cubic")
    with pytest.raises(ValueError):
        KPConv(15, 3, 8, 16, 0.05, 0.0625, aggregation_mode="mean")
    with pytest.raises(ValueError):
        block_decider("conv9", 0.1, 8, 16, 0, kpconv_config("3dmatch"))


def test_load_kernels_reads_the_reference_ply_layout(tmp_path, monkeypatch):
    """kernels/dispositions/k_015_center_3D.ply relative to the CWD wins over the generated disposition."""
    pts = np.random.default_rng(0).normal(size=(15, 3))
    pts[0] = 0
    d = tmp_path / "kernels" / "dispositions"
    d.mkdir(parents=True)
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex 15\nproperty float64 x\nproperty float64 y\n"
              "property float64 z\nend_header\n").encode()
    (d / "k_015_center_3D.ply").write_bytes(header + pts.astype("<f8").tobytes())
    monkeypatch.chdir(tmp_path)
    np.random.seed(3)
    got = load_kernels(2.0, 15, 3, "center")
    np.random.seed(3)
    theta = np.random.rand() * 2 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    rot = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
    want = np.matmul(2.0 * (pts + np.random.normal(scale=0.01, size=pts.shape)), rot).astype(np.float32)
    assert np.array_equal(got, want)


def test_segment_instance_norm_matches_torch_instance_norm():
    torch.manual_seed(0)
    lens = torch.tensor([37, 1200, 5], dtype=torch.int32)
    x = torch.randn(int(lens.sum()), 24) * 3 + 1
    got = _segment_instance_norm(x, lens)
    norm = torch.nn.InstanceNorm1d(24, momentum=0.02)
    want = torch.cat([norm(seg.t().unsqueeze(0)).squeeze(0).t() for seg in torch.split(x, lens.tolist())], 0)
    assert rel_err(got.numpy(), want.numpy()) < 1e-5


def test_synthetic_clouds_have_the_baseline_shapes():
    src, tgt, pose = synthetic.threedmatch_pair(seed=0)
    assert 15000 < len(src) < 26000 and 15000 < len(tgt) < 26000 and pose.shape == (3, 4)
    assert np.allclose(pose[:, :3] @ pose[:, :3].T, np.eye(3), atol=1e-5)
    s2, _, _ = synthetic.threedmatch_pair(seed=0)
    assert np.array_equal(src, s2)  # seeded
    src, tgt, _ = synthetic.modelnet_pair(seed=0)
    assert src.shape == (717, 3) and np.abs(src).max() <= 1.1


def test_shard_pairs_round_robin():
    shards = [shard_pairs(10, r, 4) for r in range(4)]
    assert shards == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    assert sorted(sum(shards, [])) == list(range(10))


@pytest.mark.gpu
def test_compute_overlaps_matches_reference_formula():
    """compute_overlaps (reference finegrained_kpconv.py:545-571) on a hand-made 2-level pyramid."""
    from kpreg_b200.kpconv import PreprocessorGPU, Preprocessor, compute_overlaps
    assert PreprocessorGPU is Preprocessor
    pools0 = torch.tensor([[0, 1, 5], [2, 5, 5], [3, 4, 0]]).cuda()           # 5 = pad (level 0 has 5 points)
    batch = {'src_overlap': [torch.tensor([True, False, True])], 'tgt_overlap': [torch.tensor([False, True])],
             'kpconv_meta': {'points': [torch.zeros(5, 3).cuda(), torch.zeros(3, 3).cuda()], 'pools': [pools0, torch.zeros(0, 1).cuda()],
                             'stack_lengths': [torch.tensor([3, 2]), torch.tensor([2, 1])]}}
    out = compute_overlaps(batch)
    assert torch.equal(out['pyr_0'].cpu(), torch.tensor([1., 0., 1., 0., 1.]))
    assert torch.allclose(out['pyr_1'].cpu(), torch.tensor([0.5, 1.0, (0. + 1. + 1.) / 3]))


@pytest.mark.gpu
def test_compute_overlaps_matches_reference_fixture():
    """The overlap pyramid of a 3DMatch-shape pair against the REFERENCE's compute_overlaps run on the reference's own
    pyramid (tests/golden/make_golden_r2.py): our pyramid (CUDA Preprocessor) + kpreg_overlap_pool, 1e-6 (fp32 sums of
    <= 40 terms in a different order)."""
    import os
    import numpy as np
    from conftest import GOLDEN
    from kpreg_b200 import kpconv_config, synthetic
    from kpreg_b200.kpconv import Preprocessor, compute_overlaps
    ov = dict(np.load(os.path.join(GOLDEN, "overlaps_r2.npz")))
    cfg = kpconv_config("3dmatch")
    src, tgt, _ = synthetic.threedmatch_pair(seed=2, n_raw=9000)
    for dt in (torch.int64, torch.int32):
        meta = Preprocessor(cfg, index_dtype=dt)([torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()])
        batch = {'src_overlap': [torch.from_numpy(ov["src_overlap"]).cuda()], 'tgt_overlap': [torch.from_numpy(ov["tgt_overlap"]).cuda()],
                 'kpconv_meta': meta}
        pyr = compute_overlaps(batch)
        assert sorted(pyr) == ["pyr_0", "pyr_1", "pyr_2", "pyr_3"]
        for p in range(4):
            got, want = pyr[f"pyr_{p}"].cpu().numpy(), ov[f"pyr_{p}"]
            assert got.shape == want.shape and got.dtype == np.float32
            assert float(np.abs(got - want).max()) < 1e-6, p


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["shuffled_truncated", "shuffled_all", "ordered_truncated"])
def test_shuffle_points_matches_reference(case):
    """ShufflePoints on the device against the REFERENCE class run with the same numpy seed (exact: gathers and index maps)."""
    import os
    import numpy as np
    from conftest import GOLDEN
    from kpreg_b200.ingest import ShufflePoints
    g = dict(np.load(os.path.join(GOLDEN, "shuffle_points.npz")))
    max_pts, shuffle = (int(v) for v in g[f"{case}::args"])
    data = {k.split("::")[-1]: torch.from_numpy(v).cuda() for k, v in g.items() if k.startswith(f"{case}::in::")}
    np.random.seed(5)
    out = ShufflePoints(max_pts=max_pts, shuffle=bool(shuffle))(data)
    for key in ("src_xyz", "tgt_xyz", "src_overlap", "tgt_overlap", "correspondences"):
        want = g[f"{case}::out::{key}"]
        got = out[key].cpu().numpy()
        assert got.shape == want.shape and got.dtype == want.dtype, (key, got.dtype, want.dtype)
        assert np.array_equal(got, want), key


@pytest.mark.gpu
def test_load_cloud_reads_the_reference_pth_format(tmp_path):
    import numpy as np
    from kpreg_b200.ingest import load_cloud
    cloud = np.random.default_rng(0).normal(size=(1234, 3)).astype(np.float32)
    torch.save(cloud, tmp_path / "cloud_bin_0.pth")        # the 3DMatch fragments are pickled numpy arrays
    got = load_cloud(str(tmp_path / "cloud_bin_0.pth"))
    assert got.is_cuda and got.dtype == torch.float32 and np.array_equal(got.cpu().numpy(), cloud)
