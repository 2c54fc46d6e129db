"""-m gpu: KPConv forward/backward, max-pool and the whole encoder on CUDA vs the reference's Python
(golden vectors) and vs the CPU oracle.  Tolerance (north star): features within 1e-4 relative."""
import numpy as np
import pytest
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import kpconv_config, ops, synthetic
from kpreg_b200 import kpconv_blocks
from kpreg_b200.kpconv import KPFEncoder, Preprocessor
from kpreg_b200.kpconv_blocks import KPConv, max_pool
from conftest import record_parity, rel_err
from gpu_util import LEVEL_KEYS, _levels, cuda

pytestmark = pytest.mark.gpu
TOL = 1e-4  # max|got - want| / max|want|


@pytest.mark.parametrize("infl", ["linear", "gaussian", "constant"])
@pytest.mark.parametrize("agg", ["sum", "closest"])
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("gemm", [1, 0])  # 1 = the shipped tcgen05 3xTF32 contraction, 0 = fp32 CUDA cores (backward pass)
def test_kpconv_forward_matches_reference(golden_modelnet, infl, agg, idx_dtype, gemm):
    g = golden_modelnet
    with torch.no_grad():
        out = ops.kpconv_forward(cuda(g["mn_points_1"]), cuda(g["mn_points_0"]), cuda(g["mn_pools_0"], idx_dtype),
                                 cuda(g["op_x"]), cuda(g[f"op_{infl}_{agg}_w"]), cuda(g[f"op_{infl}_{agg}_kp"]), 0.12,
                                 infl, agg, gemm=gemm)
    err = rel_err(out.cpu().numpy(), g[f"op_{infl}_{agg}_out"])
    record_parity(f"kpconv_forward[{infl},{agg},gemm={gemm}]", err, TOL)
    assert err < TOL


def test_kpconv_module_and_backward_match_reference(golden_modelnet):
    g = golden_modelnet
    np.random.seed(0)
    conv = KPConv(15, 3, 24, 40, 0.12, 0.165).cuda()
    conv.load_state_dict({"weights": torch.from_numpy(g["bw_w"]), "kernel_points": torch.from_numpy(g["bw_kp"])})
    x = cuda(g["op_x"]).requires_grad_(True)
    out = conv(cuda(g["mn_points_1"]), cuda(g["mn_points_0"]), cuda(g["mn_pools_0"]), x)
    out.backward(cuda(g["bw_g"]))
    assert rel_err(x.grad.cpu().numpy(), g["bw_dx"]) < TOL
    assert rel_err(conv.weights.grad.cpu().numpy(), g["bw_dw"]) < TOL
    assert conv.kernel_points.grad is None


def test_max_pool_forward_backward(golden_modelnet):
    g = golden_modelnet
    x = cuda(g["op_x"]).requires_grad_(True)
    out = max_pool(x, cuda(g["mn_pools_0"]))
    assert np.array_equal(out.detach().cpu().numpy(), g["op_maxpool"])  # a max: exact
    out.backward(cuda(g["bw_g"][:, :24]))
    assert rel_err(x.grad.cpu().numpy(), g["bw_maxpool_dx"]) < 1e-6


@pytest.mark.parametrize("c_in,c_out", [(1, 64), (32, 32), (64, 64), (128, 128), (256, 256), (40, 24), (30, 16), (100, 8)])
def test_kpconv_forward_matches_oracle_channel_sweep(oracle, golden_modelnet, c_in, c_out):
    """Every channel width the 3DMatch encoder uses (SURVEY.md §8a call table), + an odd one."""
    g = golden_modelnet
    rng = np.random.default_rng(c_in)
    q, s, idx = g["mn_points_0"], g["mn_points_0"], g["mn_neighbors_0"]
    x = np.ones((s.shape[0], 1), np.float32) if c_in == 1 else rng.normal(size=(s.shape[0], c_in)).astype(np.float32)
    w = (rng.normal(size=(15, c_in, c_out)) / np.sqrt(15 * c_in)).astype(np.float32)
    kp = g["op_linear_sum_kp"] * 0.5
    want = oracle.kpconv_forward(q, s, idx, x, w, kp, 0.06)
    got = ops.kpconv_forward(cuda(q), cuda(s), cuda(idx), cuda(x), cuda(w), cuda(kp), 0.06, gemm=0)
    assert rel_err(got.cpu().numpy(), want.numpy()) < TOL


@pytest.mark.parametrize("n_nbrs", [5, 32, 40, 64])
@pytest.mark.parametrize("agg,c_out", [("sum", 64), ("sum", 24), ("closest", 128)])
def test_kpconv_single_input_channel_fused_kernel(oracle, golden_modelnet, n_nbrs, agg, c_out):
    """c_in == 1 (the encoder's first block) runs gather + contraction + normalisation in one kernel: arbitrary
    feature values (negative and zero rows do not count in the normaliser), shadow indices anywhere in the row,
    widths that exercise the full-round, masked-round and short-tail (H = 40) paths."""
    g = golden_modelnet
    rng = np.random.default_rng(n_nbrs + c_out)
    q = s = g["mn_points_0"]
    n = s.shape[0]
    near = g["mn_neighbors_0"]
    idx = rng.integers(0, n + 1, size=(n, n_nbrs))                      # n == shadow
    w0 = min(near.shape[1], n_nbrs)
    idx[:, :w0] = near[:, :w0]                                          # real neighbours first: non-zero influences
    x = rng.normal(size=(n, 1)).astype(np.float32)
    x[::5] = 0.0
    w = (rng.normal(size=(15, 1, c_out)) / 4).astype(np.float32)
    kp = g["op_linear_sum_kp"] * 0.5
    want = oracle.kpconv_forward(q, s, idx, x, w, kp, 0.06, aggregation_mode=agg)
    for dt in (torch.int64, torch.int32):
        got = ops.kpconv_forward(cuda(q), cuda(s), cuda(idx, dt), cuda(x), cuda(w), cuda(kp), 0.06, aggregation=agg, gemm=1)
        assert rel_err(got.cpu().numpy(), want.numpy()) < TOL


def test_normalisation_counts_positive_feature_sums(oracle, golden_modelnet):
    """The fork divides by the number of neighbours whose feature SUM is positive (blocks :396-399)."""
    g = golden_modelnet
    rng = np.random.default_rng(5)
    q, s, idx = g["mn_points_0"], g["mn_points_0"], g["mn_neighbors_0"]
    x = rng.normal(size=(s.shape[0], 8)).astype(np.float32)
    x[::3] = -np.abs(x[::3])  # a third of the rows have a negative sum -> not counted
    w = rng.normal(size=(15, 8, 16)).astype(np.float32)
    kp = g["op_linear_sum_kp"] * 0.5
    want = oracle.kpconv_forward(q, s, idx, x, w, kp, 0.06)
    got = ops.kpconv_forward(cuda(q), cuda(s), cuda(idx), cuda(x), cuda(w), cuda(kp), 0.06)
    assert rel_err(got.cpu().numpy(), want.numpy()) < TOL


def _golden_encoder(g, cfg):
    np.random.seed(0)
    enc = KPFEncoder(cfg, 256)
    sd = {k[len("mn_sd::"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("mn_sd::")}
    enc.load_state_dict(sd, strict=True)  # reference parameter names load unchanged
    return enc.cuda()


def test_encoder_matches_reference_eval(golden_modelnet):
    g = golden_modelnet
    cfg = kpconv_config("modelnet", first_feats_dim=64)
    enc = _golden_encoder(g, cfg).eval()
    batch = {key: [cuda(a) for a in _levels(g, "mn_", key)] for key in LEVEL_KEYS}
    x0 = torch.ones((g["mn_points_0"].shape[0], 1), device="cuda")
    with torch.no_grad():
        y, skips = enc(x0, batch)
    assert rel_err(y.cpu().numpy(), g["mn_enc_out"]) < TOL
    assert len(skips) == 2 and rel_err(skips[1].cpu().numpy(), g["mn_enc_skip_1"]) < TOL


def test_encoder_end_to_end_from_raw_clouds(golden_modelnet):
    """Preprocessor (CUDA) -> KPFEncoder (CUDA) from the raw pair reproduces the reference's features."""
    g = golden_modelnet
    cfg = kpconv_config("modelnet", first_feats_dim=64)
    enc = _golden_encoder(g, cfg).eval()
    meta = Preprocessor(cfg)([cuda(g["mn_src"]), cuda(g["mn_tgt"])])
    x0 = torch.ones((meta["points"][0].shape[0], 1), device="cuda")
    with torch.no_grad():
        y, _ = enc(x0, meta)
    assert rel_err(y.cpu().numpy(), g["mn_enc_out"]) < TOL


def test_encoder_training_step_gradients(golden_modelnet):
    """forward + backward through every KPConv / max_pool of the encoder (BASELINE config 4 in small):
    loss = sum(encoder output), BatchNorm in training mode, gradients vs the reference's autograd."""
    g = golden_modelnet
    cfg = kpconv_config("modelnet", first_feats_dim=64)
    enc = _golden_encoder(g, cfg).train()
    batch = {key: [cuda(a) for a in _levels(g, "mn_", key)] for key in LEVEL_KEYS}
    x0 = torch.ones((g["mn_points_0"].shape[0], 1), device="cuda")
    y, _ = enc(x0, batch)
    y.sum().backward()
    e_fwd = rel_err(y.detach().cpu().numpy(), g["mn_enc_out_train"])
    record_parity("encoder_modelnet_train_forward_vs_reference", e_fwd, TOL)
    record_parity("encoder_modelnet_train_grad_kp1_vs_reference", rel_err(enc.encoder_blocks[1].KPConv.weights.grad.cpu().numpy(), g["mn_grad_kp1"]), 1e-3)
    record_parity("encoder_modelnet_train_grad_kp0_vs_reference", rel_err(enc.encoder_blocks[0].KPConv.weights.grad.cpu().numpy(), g["mn_grad_kp0"]), 1e-3)
    assert e_fwd < TOL
    assert rel_err(enc.encoder_blocks[1].KPConv.weights.grad.cpu().numpy(), g["mn_grad_kp1"]) < 1e-3
    assert rel_err(enc.encoder_blocks[0].KPConv.weights.grad.cpu().numpy(), g["mn_grad_kp0"]) < 1e-3


def test_encoder_3dmatch_full_size_vs_oracle(oracle):
    """BASELINE config 2 shape (one pair, ~20k points per cloud), random-init weights."""
    cfg = kpconv_config("3dmatch")
    torch.manual_seed(1)
    np.random.seed(1)
    enc = KPFEncoder(cfg, cfg.d_embed).eval()
    src, tgt, _ = synthetic.threedmatch_pair(seed=31, n_raw=20000)
    meta = Preprocessor(cfg)([cuda(src), cuda(tgt)])
    x0 = torch.ones((meta["points"][0].shape[0], 1))
    want, _ = oracle.encoder_forward(enc.state_dict(), cfg, x0, {k: [t.cpu() for t in v] for k, v in meta.items()})
    enc = enc.cuda()
    with torch.no_grad():
        got, _ = enc(x0.cuda(), meta)
    assert got.shape[1] == 1024
    err = rel_err(got.cpu().numpy(), want.numpy())
    record_parity("encoder_3dmatch_full_size_vs_oracle", err, TOL)
    assert err < TOL


def test_training_step_3dmatch_shape_vs_oracle_autograd(oracle):
    """BASELINE config 4 in small: forward + backward through all 11 KPConv ops (and the 3 max_pools) of the
    3DMatch encoder, loss = <output, fixed random probe>; weight gradients vs the CPU oracle's autograd (2e-3
    relative: the oracle itself runs fp32 with a different summation order through 11 blocks)."""
    cfg = kpconv_config("3dmatch")
    torch.manual_seed(2)
    np.random.seed(2)
    enc = KPFEncoder(cfg, cfg.d_embed).train()
    src, tgt, _ = synthetic.threedmatch_pair(seed=41, n_raw=6000)
    meta = Preprocessor(cfg)([cuda(src), cuda(tgt)])
    x0 = torch.ones((meta["points"][0].shape[0], 1))
    watch = ["encoder_blocks.0.KPConv.weights", "encoder_blocks.5.KPConv.weights", "encoder_blocks.10.KPConv.weights"]
    cpu_meta = {k: [t.cpu() for t in v] for k, v in meta.items()}

    def oracle_grads(dtype):
        sd = {k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone())
              for k, v in enc.state_dict().items()}
        for k in watch:
            sd[k].requires_grad_(True)
        with torch.enable_grad():
            out, _ = oracle.encoder_forward(sd, cfg, x0, cpu_meta, training=True, keep_grad=True, dtype=dtype)
            # a generic linear functional of the output: sum(out) alone is degenerate (the sum over a cloud of an
            # instance-normalised feature is identically zero, so its gradient is pure cancellation noise)
            probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
            (out * probe.to(dtype)).sum().backward()
        return out.detach(), {k: sd[k].grad for k in watch}, probe

    want, grads32, probe = oracle_grads(torch.float32)   # the reference's arithmetic
    _, grads64, _ = oracle_grads(torch.float64)           # ground truth, to calibrate fp32 noise through 11 blocks
    enc = enc.cuda()
    got, _ = enc(x0.cuda(), meta)
    (got * probe.cuda()).sum().backward()
    e_fwd = rel_err(got.detach().cpu().numpy(), want.detach().numpy())
    params = dict(enc.named_parameters())
    errs = {k: rel_err(params[k].grad.cpu().numpy(), grads64[k].numpy()) for k in watch}
    noise = {k: rel_err(grads32[k].numpy(), grads64[k].numpy()) for k in watch}
    print("training step: forward err", e_fwd, "| CUDA grads vs fp64 truth", errs, "| fp32 oracle vs fp64 truth", noise)
    record_parity("training_step_3dmatch_forward_vs_oracle", e_fwd, TOL)
    for k in watch:
        record_parity(f"training_step_3dmatch_grad[{k}]_vs_fp64", errs[k], 5.0 * noise[k] + 2e-3)
    assert e_fwd < TOL
    for k in watch:
        # gradients through 11 blocks of train-mode normalisation are ill-conditioned in fp32: hold the CUDA path to
        # the accuracy the reference's own fp32 arithmetic achieves against the fp64 truth
        assert errs[k] < 5.0 * noise[k] + 2e-3, (k, errs, noise)


def test_pyramid_and_encoder_mcd_full_size(oracle):
    """BASELINE config 3 shape: 2 x 120 k LiDAR-like points (sparse grid, variable row widths)."""
    cfg = kpconv_config("mcd")
    src, tgt, _ = synthetic.mcd_pair(seed=2)
    meta = Preprocessor(cfg, index_dtype=torch.int32)([cuda(src), cuda(tgt)])
    want = oracle.preprocess([src, tgt], cfg, impl="ref" if oracle.have_ref() else "port")
    from gpu_util import check_pyramid, meta_to_numpy
    check_pyramid(oracle, meta_to_numpy(meta), want, cfg)
    widths = [int(t.shape[1]) for t in meta["neighbors"]]
    assert widths[0] < 40  # sparse LiDAR: the first levels do not saturate the neighbourhood limit
    # encoder features on the same pyramid (240 k points) vs the CPU oracle
    torch.manual_seed(4)
    np.random.seed(4)
    enc = KPFEncoder(cfg, cfg.d_embed).eval()
    x0 = torch.ones((meta["points"][0].shape[0], 1))
    keys = ("points", "neighbors", "pools", "upsamples", "stack_lengths")
    ref, _ = oracle.encoder_forward(enc.state_dict(), cfg, x0, {k: [t.cpu() for t in meta[k]] for k in keys})
    enc = enc.cuda()
    with torch.no_grad():
        got, _ = enc(x0.cuda(), meta)
    err = rel_err(got.cpu().numpy(), ref.numpy())
    record_parity("encoder_mcd_full_size_vs_oracle", err, TOL)
    assert err < TOL


@pytest.mark.parametrize("c_in,c_out,n_sub", [(32, 32, None), (64, 64, None), (128, 128, 700), (256, 256, 500), (40, 24, None)])
def test_kpconv_backward_matches_oracle_autograd_channel_sweep(oracle, golden_modelnet, c_in, c_out, n_sub):
    """d_x and d_weights for every channel width of the encoder vs autograd through the CPU oracle."""
    g = golden_modelnet
    rng = np.random.default_rng(c_in + 3)
    q, s, idx = g["mn_points_1"], g["mn_points_0"], g["mn_pools_0"]
    if n_sub:
        q, idx = q[:n_sub], idx[:n_sub]
    x = rng.normal(size=(s.shape[0], c_in)).astype(np.float32)
    w = (rng.normal(size=(15, c_in, c_out)) / np.sqrt(15 * c_in)).astype(np.float32)
    kp = g["op_linear_sum_kp"]
    go = rng.normal(size=(q.shape[0], c_out)).astype(np.float32)
    xt = torch.from_numpy(x).requires_grad_(True)
    wt = torch.from_numpy(w).requires_grad_(True)
    oracle.kpconv_forward(q, s, idx, xt, wt, kp, 0.12).backward(torch.from_numpy(go))
    d_x, d_w = ops.kpconv_backward(cuda(q), cuda(s), cuda(idx), cuda(x), cuda(w), cuda(kp), cuda(go), 0.12)
    assert rel_err(d_x.cpu().numpy(), xt.grad.numpy()) < TOL
    assert rel_err(d_w.cpu().numpy(), wt.grad.numpy()) < TOL


@pytest.mark.parametrize("prefix", ["tdm", "mn128"])
def test_encoder_fused_path_matches_reference_r2_fixture(golden_encoder_r2, prefix):
    """The fused inference path (tcgen05 Linear+BN GEMMs, the register-resident res2net chain, K-concatenated
    conv3 + downsample, shortcut epilogues) against the REFERENCE's KPFEncoder run on the same clouds and weights
    (tests/golden/make_golden_r2.py): shipped 3DMatch configuration (res2net widths 28 / 56 / 112 / 224) and the
    ModelNet architecture at width 28.  Bound: the north star's 1e-4."""
    from test_oracle import r2_case, r2_state_dict
    g = golden_encoder_r2
    cfg, d_bottle, clouds = r2_case(g, prefix)
    np.random.seed(0)
    enc = KPFEncoder(cfg, d_bottle)
    enc.load_state_dict(r2_state_dict(g, prefix, enc), strict=True)
    enc = enc.cuda().eval()
    assert enc.encoder_blocks[1].res2net.layer1[0].width % 4 == 0  # the fused branch, not the stock-PyTorch one
    meta = Preprocessor(cfg, index_dtype=torch.int32)([cuda(c) for c in clouds])
    x0 = torch.ones((meta["points"][0].shape[0], 1), device="cuda")
    launches0 = kpreg_b200._lib.launch_count()
    with torch.no_grad():
        y, skips = enc(x0, meta)
    assert kpreg_b200._lib.launch_count() - launches0 > 5 * len(enc.encoder_blocks)  # the CUDA library did the work
    err = rel_err(y.cpu().numpy(), g[f"{prefix}_enc_out"])
    stride = int(g[f"{prefix}_row_stride"])
    errs = [rel_err(s.cpu().numpy()[::stride], g[f"{prefix}_skip_{i}_rows"]) for i, s in enumerate(skips)]
    print(f"fused encoder vs reference KPFEncoder ({prefix}): out {err:.2e}, skips {['%.1e' % e for e in errs]}")
    record_parity(f"encoder_fused_vs_reference_fixture[{prefix}]", err, TOL)
    for i, e in enumerate(errs):
        record_parity(f"encoder_fused_vs_reference_fixture[{prefix}].skip{i}", e, TOL)
    assert err < TOL and max(errs) < TOL


@pytest.mark.parametrize("variant", ["no_front", "no_front_no_pair", "no_front_no_chain", "no_pair_conv3", "no_shared_unary"])
def test_encoder_fallback_paths_match_reference_r2_fixture(golden_encoder_r2, variant, monkeypatch):
    """The slower routes of the fused res2net unit — conv1 as its own GEMM + the mma.sync chain kernel (no front kernel), the
    side-output chain scheme (no pair GEMM) and layer-by-layer chains at every width (no chain kernel) — against the same
    reference fixture and bound as the default route: whichever route a shape or an environment switch selects, parity holds."""
    from kpreg_b200 import kpconv_blocks as kb
    from test_oracle import r2_case, r2_state_dict
    g = golden_encoder_r2
    cfg, d_bottle, clouds = r2_case(g, "tdm")
    if variant == "no_shared_unary":
        # unary1 and the shortcut's unary as two GEMMs over the block input
        monkeypatch.setattr(kb, "SHARED_UNARY", False)
    elif variant == "no_pair_conv3":
        # wide units as before: x copied behind the concatenation, conv1's output in its own buffer, last group copied into z
        monkeypatch.setattr(kb, "PAIR_CONV3", False)
    else:
        monkeypatch.setattr(kb, "FRONT_KERNEL", False)
    if variant == "no_front_no_pair":
        monkeypatch.setattr(ops, "linear_pair_supported", lambda *a, **k: False)
    if variant == "no_front_no_chain":
        monkeypatch.setattr(kb, "CHAIN_KERNEL", False)
    np.random.seed(0)
    enc = KPFEncoder(cfg, d_bottle)
    enc.load_state_dict(r2_state_dict(g, "tdm", enc), strict=True)
    enc = enc.cuda().eval()
    meta = Preprocessor(cfg, index_dtype=torch.int32)([cuda(c) for c in clouds])
    x0 = torch.ones((meta["points"][0].shape[0], 1), device="cuda")
    with torch.no_grad():
        y, skips = enc(x0, meta)
    err = rel_err(y.cpu().numpy(), g["tdm_enc_out"])
    record_parity(f"encoder_fallback[{variant}]_vs_reference_fixture", err, TOL)
    assert err < TOL


def test_row_positive_predicate_never_flips_on_encoder_features(oracle):
    """KPConv normalises by the number of neighbours whose feature SUM is positive (reference blocks :396-399, an fp32
    torch.sum).  The CUDA kernel decides the sign from an fp64 sum; the two can only differ for rows whose sum lies
    within fp32 rounding of zero.  On the real inputs of every KPConv of the 3DMatch encoder: count the rows where the
    sign of an fp32 sum (torch.sum, either device) differs from the fp64 sign, and record how close to zero any row comes."""
    cfg = kpconv_config("3dmatch")
    torch.manual_seed(1)
    np.random.seed(1)
    enc = KPFEncoder(cfg, cfg.d_embed).eval().cuda()
    src, tgt, _ = synthetic.threedmatch_pair(seed=31, n_raw=20000)
    meta = Preprocessor(cfg, index_dtype=torch.int32)([cuda(src), cuda(tgt)])
    seen = []
    hooks = [m.register_forward_pre_hook(lambda mod, args: seen.append(args[3])) for m in enc.modules() if isinstance(m, KPConv)]
    with torch.no_grad():
        enc(torch.ones((meta["points"][0].shape[0], 1), device="cuda"), meta)
    for h in hooks:
        h.remove()
    assert len(seen) == 11
    flips, closest, rows = 0, float("inf"), 0
    for x in seen:
        s64 = x.double().sum(1)
        for s32 in (x.sum(1), x.cpu().sum(1).cuda()):            # fp32 sums in two different orders
            flips += int(((s32 > 0) != (s64 > 0)).sum())
        scale = x.abs().double().sum(1)
        nz = scale > 0
        if bool(nz.any()):
            closest = min(closest, float((s64[nz].abs() / scale[nz]).min()))
        rows += x.shape[0]
    print(f"row predicate: {rows} rows over 11 KPConv inputs, sign flips {flips}, closest |sum|/sum|x| = {closest:.2e}")
    record_parity("row_positive_predicate_sign_flips", flips, 0.5)
    record_parity("row_positive_predicate_closest_relative_sum", closest, float("inf"))
    assert flips == 0


def test_fused_caches_follow_parameter_updates():
    """The fused inference path caches derived weights (split TF32 operands, folded Linear+BatchNorm, chain packs).
    Updates through the autograd API are seen through the version counters; writes through ``.data`` need
    ``kpreg_b200.invalidate_caches()`` (documented in ops.py) — after it the fused path and stock PyTorch agree again."""
    from kpreg_b200.res2net import my_Bottle2neck, my_res2Net
    torch.manual_seed(3)
    unit = my_res2Net(my_Bottle2neck, 32, 128, baseWidth=14, scale=8).cuda().eval()
    x = torch.randn(5000, 32, device="cuda")

    def both():
        with torch.no_grad():
            fused = unit(x)
            kpconv_blocks.FUSED_GLUE = False
            try:
                stock = unit(x)
            finally:
                kpconv_blocks.FUSED_GLUE = True
        return rel_err(fused.cpu().numpy(), stock.cpu().numpy())

    assert both() < TOL
    blk = unit.layer1[0]
    with torch.no_grad():
        blk.conv1.weight.mul_(1.5)                 # bumps the version counter: picked up without help
        blk.bns[2].running_var.add_(0.3)
    assert both() < TOL
    blk.conv3.weight.data.mul_(0.5)                # .data writes bypass the counters ...
    blk.bn1.running_mean.data.add_(0.25)
    kpreg_b200.invalidate_caches()                 # ... so the caches are dropped explicitly
    assert both() < TOL
    unit.load_state_dict({k: v * 0.9 if v.is_floating_point() else v for k, v in unit.state_dict().items()})
    assert both() < TOL
