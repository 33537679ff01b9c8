"""Host planning logic (tap tables, sub-pixel phases, weight pack maps) against torch conv ops."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mmdyn_b200 import plan
from tests import emul



@pytest.fixture(autouse=True)
def _fp64_default():
    torch.manual_seed(0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().double().numpy()


def check_layer(lp, w, x, y_ref, dy, stride, pad, transposed, cin_pad=None, border=0):
    n = x.shape[0]
    flat = w.detach().double().reshape(-1).numpy()
    # forward
    if lp.fwd is not None:
        out = emul.igemm(lp.fwd, nhwc(x), emul.pack(flat, lp.idx_fwd), n)
        if lp.fwd.out_mode == 3:
            np.testing.assert_allclose(out, y_ref.double().numpy(), atol=1e-9)
        else:
            np.testing.assert_allclose(out[..., :y_ref.shape[1]], nhwc(y_ref), atol=1e-9)
    # reference grads
    xr = x.clone().double().requires_grad_(True)
    wr = w.clone().double().requires_grad_(True)
    op = F.conv_transpose2d if transposed else F.conv2d
    yr = op(xr, wr, stride=stride, padding=pad)
    yr.backward(dy.double())
    dyn = nhwc(dy)
    if cin_pad:
        dyn = np.concatenate([dyn, np.zeros(dyn.shape[:-1] + (cin_pad - dyn.shape[-1],))], -1)
    if border:  # zero border of `border` pixels around every image (overlapping-window operand layout)
        dyn = np.pad(dyn, ((0, 0), (border, border), (border, border), (0, 0)))
    dx = emul.igemm(lp.dgrad, dyn, emul.pack(flat, lp.idx_dgrad), n)
    np.testing.assert_allclose(dx[..., :x.shape[1]], nhwc(xr.grad), atol=1e-9)
    # weight gradient
    if transposed:
        G, Nat = dyn, nhwc(x).reshape(n, -1, x.shape[1])
    else:
        G, Nat = nhwc(x), dyn.reshape(n, -1, dyn.shape[-1])
    dWp = emul.wgrad(lp.wgrad, G, Nat, n)
    dW = emul.unpack_add(dWp, lp.idx_wgrad, flat.size)
    np.testing.assert_allclose(dW, wr.grad.reshape(-1).numpy(), atol=1e-8)


@pytest.mark.parametrize("Cin,Cout,H", [(32, 64, 8), (64, 128, 4)])
def test_conv_s2(Cin, Cout, H):
    w = torch.randn(Cout, Cin, 4, 4)
    x = torch.randn(2, Cin, H, H)
    y = F.conv2d(x, w, stride=2, padding=1)
    check_layer(plan.conv_s2_plan("c", 0, Cin, Cout, H), w, x, y, torch.randn_like(y), 2, 1, False)


def test_conv_k4s1p0():
    w = torch.randn(256, 128, 4, 4)
    x = torch.randn(2, 128, 8, 8)
    y = F.conv2d(x, w)
    check_layer(plan.conv_k4s1p0_plan("c4", 0, 128, 256, 8), w, x, y, torch.randn_like(y), 1, 0, False)


def test_deconv_k4s1p0():
    w = torch.randn(256, 128, 4, 4)
    x = torch.randn(2, 256, 5, 5)
    y = F.conv_transpose2d(x, w)
    check_layer(plan.deconv_k4s1p0_plan("d1", 0, 256, 128, 5), w, x, y, torch.randn_like(y), 1, 0, True)


@pytest.mark.parametrize("Cin,Cout,H", [(128, 64, 4), (64, 32, 8)])
def test_deconv_s2(Cin, Cout, H):
    w = torch.randn(Cin, Cout, 4, 4)
    x = torch.randn(2, Cin, H, H)
    y = F.conv_transpose2d(x, w, stride=2, padding=1)
    check_layer(plan.deconv_s2_plan("d", 0, Cin, Cout, H), w, x, y, torch.randn_like(y), 2, 1, True)


def test_deconv_s2_pixel_pair_gradients():
    """deconv3's backward over 128-byte pixel pairs of the x-padded gradient: same packed weights, same results"""
    Cin, Cout, H = 64, 32, 16
    w = torch.randn(Cin, Cout, 4, 4).requires_grad_(True)
    x = torch.randn(2, Cin, H, H).requires_grad_(True)
    y = F.conv_transpose2d(x, w, stride=2, padding=1)
    dy = torch.randn_like(y)
    y.backward(dy)
    lp = plan.deconv_s2_plan("d3", 0, Cin, Cout, H)
    flat = w.detach().double().reshape(-1).numpy()
    g = np.zeros((2, 2 * H, 2 * H + 2, Cout))
    g[:, :, 1:-1] = nhwc(dy)
    dx = emul.igemm(lp.extra["pair_dgrad"], g, emul.pack(flat, lp.idx_dgrad), 2)
    np.testing.assert_allclose(dx, nhwc(x.grad), atol=1e-9)
    dWp = emul.wgrad(lp.extra["pair_wgrad"], g, nhwc(x.detach()).reshape(2, -1, Cin), 2)
    dW = emul.unpack_add(dWp, lp.idx_wgrad, flat.size)
    np.testing.assert_allclose(dW, w.grad.reshape(-1).numpy(), atol=1e-8)


def test_deconv_out():
    w = torch.randn(32, 3, 4, 4)
    x = torch.randn(2, 32, 8, 8)
    y = F.conv_transpose2d(x, w, stride=2, padding=1)
    check_layer(plan.deconv_out_plan("d4", 0, 32, 3, 8), w, x, y, torch.randn_like(y), 2, 1, True, cin_pad=plan.LOGIT_CP, border=1)


def test_conv1_wgrad_window_form():
    """vae.py:198 Conv2d(3, 32, 4, 2, 1): the weight gradient read through 4-pixel windows of the padded NHWC4 input"""
    w = torch.randn(32, 3, 4, 4).requires_grad_(True)
    x = torch.randn(2, 3, 64, 64)
    y = F.conv2d(x, w, stride=2, padding=1)
    dy = torch.randn_like(y)
    y.backward(dy)
    wg, idx = plan.conv1_wgrad_plan(0)
    xp = np.zeros((2, 66, 66, plan.LOGIT_CP))
    xp[:, 1:65, 1:65, :3] = nhwc(x)
    dWp = emul.wgrad(wg, xp, nhwc(dy).reshape(2, -1, 32), 2)
    dW = emul.unpack_add(dWp, idx, w.numel())
    np.testing.assert_allclose(dW, w.grad.reshape(-1).numpy(), atol=1e-8)


def test_linear_permuted_and_concat():
    K, N1, N2 = 128, 64, 64
    w1, w2 = torch.randn(N1, K), torch.randn(N2, K)
    b1, b2 = torch.randn(N1), torch.randn(N2)
    flat = torch.cat([w1.reshape(-1), b1, w2.reshape(-1), b2]).double().numpy()
    offs = [0, N1 * K, N1 * K + N1, N1 * K + N1 + N2 * K]
    kperm = np.random.RandomState(0).permutation(K)
    lp = plan.linear_plan("heads", [offs[0], offs[2]], [offs[1], offs[3]], K, [N1, N2], k_perm=kperm)
    x = torch.randn(5, K)
    xp = x[:, kperm].double().numpy().reshape(5, 1, 1, K)
    out = emul.igemm(lp.fwd, xp, emul.pack(flat, lp.idx_fwd), 5, bias=flat[lp.bias_idx])
    ref = torch.cat([F.linear(x, w1, b1), F.linear(x, w2, b2)], 1).double().numpy()
    np.testing.assert_allclose(out.reshape(5, -1), ref, atol=1e-9)
    dy = torch.randn(5, N1 + N2).double().numpy()
    dx = emul.igemm(lp.dgrad, dy.reshape(5, 1, 1, -1), emul.pack(flat, lp.idx_dgrad), 5).reshape(5, K)
    dx_ref = dy @ torch.cat([w1, w2], 0).double().numpy()
    np.testing.assert_allclose(dx, dx_ref[:, kperm], atol=1e-9)


def test_nhwc_perm_matches_flatten():
    x = torch.randn(2, 256, 5, 5)
    perm = plan.nhwc_perm(256, 5, 5)
    a = x.permute(0, 2, 3, 1).reshape(2, -1)
    b = x.reshape(2, -1)[:, perm]
    assert torch.equal(a, b)


def test_heuristics():
    lp = plan.linear_plan("fc", [0], [6400 * 512], 6400, [512])
    assert plan.choose_ksplit(lp.fwd, 128) > 1
    assert plan.choose_ksplit(lp.fwd, 128 * 200) == 1
    c2 = plan.conv_s2_plan("c2", 0, 32, 64, 32)
    assert plan.choose_row_splits(c2.wgrad, 128) >= 1


def test_linear_plan_with_condition_columns_partitions_the_weight():
    """CVAE (vae.py:196, 257): Linear(K + cd, N) weights keep the reference's shape; the tensor-core packing
    indexes columns 0..K-1 at row pitch K + cd, the rank-cd fp32 kernels own the remaining columns.  Together
    they touch every weight element exactly once."""
    import numpy as np
    from mmdyn_b200 import plan
    K, N, cd, w_off, b_off = 512, 256, 3, 1000, 900000
    lp = plan.linear_plan("heads", [w_off, w_off + N * (K + cd)], [b_off, b_off + N], K, [N, N], ld=K + cd)
    idx = lp.idx_fwd
    assert idx.shape == (2 * N, K) and lp.fwd.N == 2 * N and lp.fwd.K == K
    rows = (idx - w_off) // (K + cd)
    cols = (idx - w_off) % (K + cd)
    assert (cols < K).all() and (rows == np.arange(2 * N)[:, None]).all()
    packed = set(idx.reshape(-1).tolist())
    cond = {w_off + n * (K + cd) + K + j for n in range(2 * N) for j in range(cd)}
    assert not (packed & cond) and len(packed | cond) == 2 * N * (K + cd)
    # the decoder's upsample Linear(256 + cd, 6400) with the NHWC output permutation
    up = plan.linear_plan("up", [0], [10 ** 7], 256, [6400], n_perm=plan.nhwc_perm(256, 5, 5), ld=256 + cd)
    assert up.idx_fwd.shape == (6400, 256) and (up.idx_fwd % (256 + cd) < 256).all()
    assert sorted((up.idx_fwd[:, 0] // (256 + cd)).tolist()) == list(range(6400))  # a permutation of the rows
