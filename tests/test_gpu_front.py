"""kpreg_front_forward (conv1 + res2net chain in one tcgen05 kernel) against an fp64 evaluation of the definition
(reference models/backbone_kpconv/res2net.py:125-152 with the eval-mode BatchNorm folded in)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-4  # north_star feature tolerance (relative to the output's scale)


def _reference(x, w1, b1, wc, bc, width, n_groups):
    x, w1, b1, wc, bc = (a.double() for a in (x, w1, b1, wc, bc))
    t = torch.relu(x @ w1.t() + b1)
    groups = torch.split(t, width, 1)
    outs, carry = [], None
    for g in range(n_groups - 1):
        carry = groups[g] if g == 0 else carry + groups[g]
        carry = torch.relu(carry @ wc[g].t() + bc[g])
        outs.append(carry)
    outs.append(groups[n_groups - 1])
    return torch.cat(outs, 1)


@pytest.mark.parametrize("width,c_in,n_groups", [(28, 32, 8), (56, 64, 8), (32, 32, 8), (64, 64, 8), (16, 8, 2), (28, 24, 5),
                                                   (40, 64, 3), (56, 48, 8)])
@pytest.mark.parametrize("m", [1, 127, 128, 129, 1000, 70001])
@pytest.mark.parametrize("copy_x", [False, True])
def test_front_forward_vs_fp64(width, c_in, n_groups, m, copy_x):
    from kpreg_b200 import ops
    assert ops.front_supported(width, n_groups, c_in)
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(1000 * width + 10 * c_in + n_groups + m)
    x = torch.randn(m, c_in, generator=g).to(dev)
    w1 = (torch.randn(n_groups * width, c_in, generator=g) / c_in ** 0.5).to(dev)
    b1 = (0.3 * torch.randn(n_groups * width, generator=g)).to(dev)
    wc = (torch.randn(n_groups - 1, width, width, generator=g) / width ** 0.5).to(dev)
    bc = (0.3 * torch.randn(n_groups - 1, width, generator=g)).to(dev)
    pack = ops.FrontPack(w1, b1, wc, bc)
    k_cat = n_groups * width
    z = torch.full((m, k_cat + (c_in if copy_x else 0) + 4), -7.0, device=dev)  # 4 guard columns the kernel must not touch
    ops.front_forward(x, pack, z, copy_x=copy_x)
    torch.cuda.synchronize()
    ref = _reference(x, w1, b1, wc, bc, width, n_groups)
    err = (z[:, :k_cat].double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
    assert err <= TOL, f"relative error {err:.3e}"
    if copy_x:
        assert torch.equal(z[:, k_cat:k_cat + c_in], x)
        assert (z[:, k_cat + c_in:] == -7.0).all()
    else:
        assert (z[:, k_cat:] == -7.0).all()


def test_front_unsupported_shapes():
    from kpreg_b200 import ops
    assert not ops.front_supported(112, 8, 128)
    assert not ops.front_supported(30, 8, 32)   # width % 4
    assert not ops.front_supported(28, 9, 32)
    assert not ops.front_supported(28, 8, 64)   # narrow groups with a wide input: not instantiated
