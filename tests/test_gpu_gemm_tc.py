"""-m gpu: the tcgen05 (3xTF32) contraction vs the fp32 CUDA-core contraction and the CPU oracle.
The split-operand scheme must hold the same 1e-4 relative tolerance as the fp32 path."""
import numpy as np
import pytest
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200 import ops
from conftest import rel_err
from gpu_util import cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c_in,c_out", [(32, 32), (64, 64), (128, 128), (256, 256), (24, 40), (32, 28), (12, 200)])
def test_tc_contraction_matches_fp32_and_oracle(oracle, golden_modelnet, c_in, c_out):
    g = golden_modelnet
    rng = np.random.default_rng(c_in * 7 + c_out)
    q, s, idx = g["mn_points_0"], g["mn_points_0"], g["mn_neighbors_0"]
    x = rng.normal(size=(s.shape[0], c_in)).astype(np.float32)
    w = (rng.normal(size=(15, c_in, c_out)) / np.sqrt(15 * c_in)).astype(np.float32)
    kp = g["op_linear_sum_kp"] * 0.5
    args = (cuda(q), cuda(s), cuda(idx), cuda(x), cuda(w), cuda(kp), 0.06)
    simt = ops.kpconv_forward(*args, gemm=0)
    tc = ops.kpconv_forward(*args, gemm=1)
    torch.cuda.synchronize()
    want = oracle.kpconv_forward(q, s, idx, x, w, kp, 0.06, dtype=torch.float64)
    e_tc, e_simt = rel_err(tc.cpu().numpy(), want.numpy()), rel_err(simt.cpu().numpy(), want.numpy())
    print(f"c_in={c_in} c_out={c_out}: err vs fp64 oracle  tcgen05 3xTF32 {e_tc:.2e}   fp32 CUDA cores {e_simt:.2e}")
    assert e_tc < 5e-6      # fp32-grade, far inside the 1e-4 budget
    assert e_simt < 1e-4    # sequential fp32 accumulation over K*c_in terms


def test_tc_contraction_many_row_tiles(oracle):
    """M spanning many 128-row tiles with a ragged tail, wide K (several pipeline rounds)."""
    rng = np.random.default_rng(0)
    n = 128 * 37 + 19
    pts = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    idx = rng.integers(0, n + 1, size=(n, 9)).astype(np.int32)  # includes shadow indices (= n)
    x = rng.normal(size=(n, 64)).astype(np.float32)
    w = rng.normal(size=(15, 64, 64)).astype(np.float32)
    kp = rng.uniform(-0.3, 0.3, size=(15, 3)).astype(np.float32)
    args = (cuda(pts), cuda(pts), cuda(idx), cuda(x), cuda(w), cuda(kp), 1.5)
    simt = ops.kpconv_forward(*args, gemm=0)
    tc = ops.kpconv_forward(*args, gemm=1)
    torch.cuda.synchronize()
    want = oracle.kpconv_forward(pts, pts, idx, x, w, kp, 1.5, dtype=torch.float64)
    assert rel_err(tc.cpu().numpy(), want.numpy()) < 5e-6
    assert rel_err(simt.cpu().numpy(), want.numpy()) < 1e-4


@pytest.mark.parametrize("m,k,n", [(5000, 32, 128), (1000, 28, 28), (777, 224, 224), (4096, 128, 32), (130, 1024, 256), (300, 3, 16),
                                   (513, 40, 30), (64, 96, 72), (2000, 1100, 40), (300, 1100, 200), (1000, 2048, 130)])
@pytest.mark.parametrize("gemm", [1, 0])
def test_linear_forward_matches_torch(m, k, n, gemm):
    """nn.Linear(bias=False) + folded BatchNorm + residual + activation (+ the chained second output)."""
    torch.manual_seed(m + k + n)
    wide = torch.randn(m, 2 * k + 8, device="cuda")
    x = wide[:, 4:4 + k]  # a column slice: rows contiguous, pitch != k
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    scale, shift = torch.rand(n, device="cuda") + 0.5, torch.randn(n, device="cuda")
    res, add = torch.randn(m, n, device="cuda"), torch.randn(m, n, device="cuda")
    out_wide = torch.zeros(m, n + 8, device="cuda")
    out2 = torch.empty(m, n, device="cuda")
    got = ops.linear_forward(x, w, scale, shift, residual=res, act="relu", out=out_wide[:, 4:4 + n], out2=out2, addend=add,
                             gemm=gemm)
    want = torch.relu((x.double() @ w.double().t()) * scale.double() + shift.double() + res.double())
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < 1e-5
    assert rel_err(out2.cpu().numpy(), (want + add.double()).cpu().numpy()) < 1e-5
    assert float(out_wide[:, :4].abs().max()) == 0.0 and float(out_wide[:, 4 + n:].abs().max()) == 0.0  # no stray writes
    plain = ops.linear_forward(x, w, gemm=gemm)
    assert rel_err(plain.cpu().numpy(), (x.double() @ w.double().t()).cpu().numpy()) < 1e-5
    leaky = ops.linear_forward(x, w, act="leaky_relu", slope=0.1, gemm=gemm)
    assert rel_err(leaky.cpu().numpy(), torch.nn.functional.leaky_relu(x.double() @ w.double().t(), 0.1).cpu().numpy()) < 1e-5


@pytest.mark.parametrize("lens,c", [([37, 1200, 5], 24), ([20000, 21000, 1, 300], 64), ([129] * 16, 1024)])
def test_segment_norm_matches_instance_norm(lens, c):
    torch.manual_seed(c)
    lens_t = torch.tensor(lens, dtype=torch.int32, device="cuda")
    x = torch.randn(sum(lens), c, device="cuda") * 3 + 1
    res = torch.randn_like(x)
    norm = torch.nn.InstanceNorm1d(c)
    want = torch.cat([norm(seg.t().unsqueeze(0)).squeeze(0).t() if seg.shape[0] > 1 else torch.zeros_like(seg)
                      for seg in torch.split(x.double(), lens)], 0)
    got = ops.segment_norm(x, lens_t)
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < 1e-5
    got2 = ops.segment_norm(x, lens_t, residual=res, act="leaky_relu", slope=0.1)
    want2 = torch.nn.functional.leaky_relu(want + res.double(), 0.1)
    assert rel_err(got2.cpu().numpy(), want2.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("lens,c", [([37, 1200, 5, 3001], 32), ([20000, 21000, 1, 300], 64), ([700, 650], 128)])
def test_segment_norm_row_predicate_feeds_kpconv(lens, c):
    """segment_norm(..., row_pos=True): where the kernel supports it (32 / 64 channels) the output carries KPConv's row
    predicate (sum of the row's features > 0, finegrained_kpconv_blocks.py:396-397) — equal to the definition evaluated on the
    output, identical values with and without it, dropped once the tensor is written to — and kpconv_forward gives the same
    result with the tag as with its own pass over the features."""
    torch.manual_seed(c + 1)
    lens_t = torch.tensor(lens, dtype=torch.int32, device="cuda")
    n = sum(lens)
    x = torch.randn(n, c, device="cuda") * 2 + 0.05
    plain = ops.segment_norm(x, lens_t, act="leaky_relu", slope=0.1)
    tagged = ops.segment_norm(x, lens_t, act="leaky_relu", slope=0.1, row_pos=True)
    assert torch.equal(plain, tagged)
    tag = ops.row_predicate_of(tagged)
    assert ops.row_predicate_of(plain) is None
    if c > 64:
        assert tag is None
        return
    flags, ptr = tag
    assert ptr == tagged.data_ptr() and flags.dtype == torch.uint8 and flags.shape == (n,)
    want = tagged.double().sum(1).float() > 0
    assert torch.equal(flags.bool(), want)
    raw = ops.segment_norm(x, lens_t, row_pos=True)  # no activation: row sums of either sign
    want_raw = raw.double().sum(1).float() > 0
    assert torch.equal(ops.row_predicate_of(raw)[0].bool(), want_raw)
    assert 0.2 < want_raw.float().mean().item() < 0.8
    # the KPConv that consumes it
    gen = torch.Generator(device="cuda").manual_seed(5)
    pts = torch.rand((n, 3), device="cuda", generator=gen)
    idx = torch.randint(0, n + 40, (n, 12), device="cuda", generator=gen).clamp(max=n).to(torch.int32)  # some shadow entries
    weights = torch.randn((15, c, 16), device="cuda", generator=gen) / c ** 0.5
    kp = torch.rand((15, 3), device="cuda", generator=gen) * 0.2 - 0.1
    with torch.no_grad():
        a = ops.kpconv_forward(pts, pts, idx, tagged, weights, kp, 0.15, gemm=1)
        b = ops.kpconv_forward(pts, pts, idx, tagged.clone(), weights, kp, 0.15, gemm=1)
    assert torch.equal(a, b)
    tagged.add_(1.0)                       # written to: the tag no longer describes the tensor
    assert ops.row_predicate_of(tagged) is None
    with torch.inference_mode():           # inference tensors carry no version counter: never tagged, same values
        inf = ops.segment_norm(x, lens_t, act="leaky_relu", slope=0.1, row_pos=True)
        assert ops.row_predicate_of(inf) is None and torch.equal(inf, plain)
        c_inf = ops.kpconv_forward(pts, pts, idx, inf, weights, kp, 0.15, gemm=1)
    assert torch.equal(c_inf, a)


@pytest.mark.parametrize("w,n_layers,m,c_x", [(28, 7, 5000, 32), (56, 7, 3001, 64), (14, 7, 257, 16), (16, 3, 31, 0), (64, 5, 700, 30),
                                              (56, 7, 256, 0)])
def test_chain_kernel_matches_layerwise_reference(w, n_layers, m, c_x):
    """res2net's chained layers in one kernel (kpreg_chain_forward) vs the layer-by-layer definition in fp64
    (reference res2net.py:137-152): every out_i, the pass-through group and the copy of the block input."""
    torch.manual_seed(w * 100 + m)
    groups = n_layers + 1
    assert ops.chain_supported(w, n_layers)
    wide = torch.randn(m, groups * w + 6, device="cuda")
    t = wide[:, 2:2 + groups * w]                                  # a column slice: row pitch != groups * w
    weights = torch.randn(n_layers, w, w, device="cuda") / w ** 0.5
    shifts = torch.randn(n_layers, w, device="cuda") * 0.3
    x = torch.randn(m, c_x, device="cuda") if c_x else None
    z = torch.full((m, groups * w + c_x + 4), 7.0, device="cuda")  # trailing guard columns must stay untouched
    ops.chain_forward(t, ops.ChainPack(weights, shifts), z, x)
    td, want = t.double(), []
    sp = td[:, :w]
    for i in range(n_layers):
        sp = torch.relu(sp @ weights[i].double().t() + shifts[i].double())
        want.append(sp)
        sp = sp + td[:, (i + 1) * w:(i + 2) * w]
    want.append(td[:, n_layers * w:])
    if c_x:
        want.append(x.double())
    want = torch.cat(want, 1)
    got = z[:, :want.shape[1]]
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < 1e-5
    assert torch.equal(got[:, n_layers * w:], want[:, n_layers * w:].float())   # copies are exact
    assert float((z[:, want.shape[1]:] - 7.0).abs().max()) == 0.0


def test_chain_kernel_unsupported_widths_are_reported():
    assert not ops.chain_supported(27, 7)      # odd width
    assert not ops.chain_supported(112, 7)     # wider than the register-resident kernel serves
    assert not ops.chain_supported(64, 7)      # packed weights exceed shared memory


def test_bottleneck_fused_chain_matches_stock_layers():
    """my_Bottle2neck inference: chain kernel on / off / stock PyTorch layers give the same features."""
    from kpreg_b200 import kpconv_blocks
    from kpreg_b200.res2net import my_Bottle2neck, my_res2Net
    torch.manual_seed(3)
    net = my_res2Net(my_Bottle2neck, 32, 128, baseWidth=14, scale=8).cuda().eval()
    for mod in net.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.normal_(0, 0.2)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.2)
    x = torch.randn(3333, 32, device="cuda")
    with torch.no_grad():
        kpconv_blocks.FUSED_GLUE = False
        try:
            stock = net(x)
        finally:
            kpconv_blocks.FUSED_GLUE = True
        chain = net(x)
        kpconv_blocks.CHAIN_KERNEL = False
        try:
            layers = net(x)
        finally:
            kpconv_blocks.CHAIN_KERNEL = True
    assert rel_err(chain.cpu().numpy(), stock.cpu().numpy()) < 1e-5
    assert rel_err(layers.cpu().numpy(), stock.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("m,k,n", [(3000, 256, 128), (517, 60, 28), (200, 1100, 72)])
@pytest.mark.parametrize("gemm", [1, 0])
def test_linear_post_residual(m, k, n, gemm):
    """out = leaky_relu(relu(x W^T + shift) + shortcut): a block's identity shortcut riding on res2net's last GEMM."""
    torch.manual_seed(m + n)
    x, w = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda") / k ** 0.5
    shift, short = torch.randn(n, device="cuda"), torch.randn(m, n, device="cuda")
    got = ops.linear_forward(x, w, None, shift, act="relu", post_residual=short, post_act="leaky_relu", gemm=gemm)
    want = torch.nn.functional.leaky_relu(torch.relu(x.double() @ w.double().t() + shift.double()) + short.double(), 0.1)
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < 1e-5


def test_res2net_unit_with_fused_shortcut():
    from kpreg_b200 import kpconv_blocks
    from kpreg_b200.res2net import my_Bottle2neck, my_res2Net
    torch.manual_seed(4)
    net = my_res2Net(my_Bottle2neck, 64, 256, baseWidth=14, scale=8).cuda().eval()
    x, short = torch.randn(2111, 64, device="cuda"), torch.randn(2111, 256, device="cuda")
    with torch.no_grad():
        fused = net(x, short)
        kpconv_blocks.FUSED_GLUE = False
        try:
            stock = torch.nn.functional.leaky_relu(net(x) + short, 0.1)
            stock2 = net(x, short)
        finally:
            kpconv_blocks.FUSED_GLUE = True
    assert rel_err(fused.cpu().numpy(), stock.cpu().numpy()) < 1e-5
    assert torch.equal(stock, stock2)


@pytest.mark.parametrize("m,k,n", [(5000, 32, 128), (20000, 128, 32), (3001, 256, 224), (777, 64, 64), (12345, 1024, 256), (40, 16, 8)])
def test_linear_autograd_on_tensor_cores(m, k, n):
    """y = x W^T, dx = dy W and dW = dy^T x (split-k over the rows, MN-major tcgen05 operands) against fp64."""
    torch.manual_seed(m)
    x = torch.randn(m, k, device="cuda", requires_grad=True)
    lin = torch.nn.Linear(k, n, bias=False).cuda()
    g = torch.randn(m, n, device="cuda")
    y = ops.linear_train(x, lin)
    assert y.grad_fn is not None and "LinearFn" in type(y.grad_fn).__name__
    y.backward(g)
    x64, w64, g64 = x.detach().double(), lin.weight.detach().double(), g.double()
    assert rel_err(y.detach().cpu().numpy(), (x64 @ w64.t()).cpu().numpy()) < 1e-5
    assert rel_err(x.grad.cpu().numpy(), (g64 @ w64).cpu().numpy()) < 1e-5
    assert rel_err(lin.weight.grad.cpu().numpy(), (g64.t() @ x64).cpu().numpy()) < 1e-5


def test_segment_norm_backward_matches_torch_autograd():
    from kpreg_b200.kpconv_blocks import _SegNormFn, _segment_instance_norm
    torch.manual_seed(0)
    lens = torch.tensor([700, 1, 1300, 2999, 256, 3], dtype=torch.int32, device="cuda")
    n = int(lens.sum())
    for c in (32, 128, 68):
        x = (torch.randn(n, c, device="cuda") * 3 + 1).requires_grad_(True)
        g = torch.randn(n, c, device="cuda")
        y = _SegNormFn.apply(x, lens)
        y.backward(g)
        x64 = x.detach().double().requires_grad_(True)
        y64 = _segment_instance_norm(x64, lens)
        y64.backward(g.double())
        assert rel_err(y.detach().cpu().numpy(), y64.detach().cpu().numpy()) < 1e-5
        # single-point clouds have zero variance: rstd = 1/sqrt(eps) amplifies rounding there; compare the rest tightly
        keep = torch.ones(n, dtype=torch.bool)
        keep[700] = False
        assert rel_err(x.grad.cpu()[keep].numpy(), x64.grad.cpu()[keep].numpy()) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("m,k1,k2,n", [(1000, 112, 112, 112), (70001, 224, 224, 224), (129, 28, 28, 28), (5000, 56, 40, 64), (1, 8, 4, 8)])
def test_linear_pair_forward_vs_fp64(m, k1, k2, n):
    """kpreg_linear_pair_forward: one GEMM whose reduction runs over two tensors (column slices of wider buffers)."""
    from kpreg_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(m + k1 + n)
    buf1 = torch.randn(m, k1 + 2 * ((k1 + 3) // 4 * 4), generator=g).to(dev)
    buf2 = torch.randn(m, k2 + 8, generator=g).to(dev)
    x1 = buf1[:, (k1 + 3) // 4 * 4:(k1 + 3) // 4 * 4 + k1]
    x2 = buf2[:, 4:4 + k2]
    w1 = (torch.randn(n, k1, generator=g) / k1 ** 0.5).to(dev)
    w2 = (torch.randn(n, k2, generator=g) / k2 ** 0.5).to(dev)
    shift = (0.2 * torch.randn(n, generator=g)).to(dev)
    wcat = torch.cat([w1, w1.new_zeros(n, (-k1) % 32), w2], 1).contiguous()
    out_buf = torch.full((m, n + 4), -3.0, device=dev)
    ops.linear_pair_forward(x1, x2, wcat, shift, act="relu", out=out_buf[:, :n])
    ref = torch.relu(x1.double() @ w1.double().t() + x2.double() @ w2.double().t() + shift.double())
    err = (out_buf[:, :n].double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
    assert err <= 1e-4, err
    assert (out_buf[:, n:] == -3.0).all()
    # the chain form: same matrix on both inputs
    wp = ops.pair_weight(w1) if k1 == k2 else None
    if wp is not None:
        y = ops.linear_pair_forward(x1, x2, wp)
        ref2 = (x1.double() + x2.double()) @ w1.double().t()
        assert (y.double() - ref2).abs().max().item() / max(ref2.abs().max().item(), 1e-30) <= 1e-4
