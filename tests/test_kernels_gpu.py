"""Kernel-level parity on a real B200: every C-ABI entry point against a torch fp64/fp32 CPU
reference of the same op, on the layer shapes of the reference model (vae.py:197-216, 263-279).

Tolerances: the tcgen05 kernels take fp16 operands (10-bit mantissa, TF32-equivalent) and
accumulate in fp32.  References are computed in fp64 from the *same fp16-rounded operands*, so
the only differences are accumulation order and the output rounding: rel 1e-3 of the output
scale for fp16 outputs, 1e-5 for fp32 outputs."""

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from mmdyn_b200 import plan  # noqa: E402


def _ops():
    from mmdyn_b200 import ops
    return ops


DEV = "cuda"


def nhwc16(x):
    return x.permute(0, 2, 3, 1).contiguous().half().to(DEV)


def packed(w, idx):
    ops = _ops()
    flat = w.reshape(-1).float().to(DEV)
    Wp = torch.empty(idx.shape, dtype=torch.float16, device=DEV)
    ops.pack_f16(flat, torch.from_numpy(idx).to(DEV), Wp)
    return Wp


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def r16(x):
    return x.half().double()


def run_fwd(geom, idx, w, A, n, out_dtype=torch.float16, bias=None, ksplit=1, out_mode=None):
    ops = _ops()
    if geom.out_mode == 3:
        out = torch.zeros(n, 3, geom.OH, geom.OW, dtype=torch.float32, device=DEV)
    else:
        out = torch.zeros(n, geom.OH, geom.OW, geom.ldc, dtype=out_dtype, device=DEV)
    ops.igemm(geom, A, packed(w, idx), out, n, bias=bias, ksplit=ksplit, out_mode=out_mode)
    torch.cuda.synchronize()
    return out


def conv_case(lp, w, x, stride, pad, transposed, n):
    op = F.conv_transpose2d if transposed else F.conv2d
    xr = r16(x).requires_grad_(True)
    wr = r16(w).requires_grad_(True)
    y = op(xr, wr, stride=stride, padding=pad)
    dy = torch.randn_like(y).half().double()
    y.backward(dy)
    return xr, wr, y.detach(), dy


def check_conv_layer(lp, w, x, stride, pad, transposed, grad_pad=None, border=0):
    ops = _ops()
    n = x.shape[0]
    xr, wr, y, dy = conv_case(lp, w, x, stride, pad, transposed, n)
    # forward
    out = run_fwd(lp.fwd, lp.idx_fwd, w, nhwc16(x), n)
    if lp.fwd.out_mode == 3:
        assert rel_err(out, y) < 1e-5 * 50, rel_err(out, y)
    else:
        assert rel_err(out.permute(0, 3, 1, 2)[:, :y.shape[1]], y) < 1e-3
    # dgrad
    dyn = dy.float()
    if grad_pad:
        dyn = torch.cat([dyn, torch.zeros(n, grad_pad - dyn.shape[1], *dyn.shape[2:])], 1)
    if border:  # zero border: the overlapping-window operand layout of the logits layer (plan.deconv_out_plan)
        dyn = F.pad(dyn, (border, border, border, border))
    dA = nhwc16(dyn)
    dx = run_fwd(lp.dgrad, lp.idx_dgrad, w, dA, n)
    assert rel_err(dx.permute(0, 3, 1, 2)[:, :x.shape[1]], xr.grad) < 1e-3
    # wgrad
    if transposed:
        G, Nat = dA, nhwc16(x)
    else:
        G, Nat = nhwc16(x), dA
    dWp = torch.zeros(lp.wgrad.Cn, lp.wgrad.K, dtype=torch.float32, device=DEV)
    ops.wgrad(lp.wgrad, G, Nat, dWp, n, scale=0.5, row_splits=plan.choose_row_splits(lp.wgrad, n))
    dW = torch.zeros(w.numel(), dtype=torch.float32, device=DEV)
    ops.unpack_add_f32(dWp, torch.from_numpy(lp.idx_wgrad).to(DEV), dW)
    torch.cuda.synchronize()
    assert rel_err(dW, 0.5 * wr.grad.reshape(-1)) < 2e-5


def test_library_loads_and_inits():
    from mmdyn_b200 import lib
    l = lib.init(0)
    assert l.mmdyn_version() >= 100


@pytest.mark.parametrize("n", [3, 40])
def test_conv2_s2(n):
    torch.manual_seed(1)
    w = torch.randn(64, 32, 4, 4) * 0.05
    x = torch.randn(n, 32, 32, 32)
    check_conv_layer(plan.conv_s2_plan("c2", 0, 32, 64, 32), w, x, 2, 1, False)


def test_conv3_s2():
    torch.manual_seed(2)
    w = torch.randn(128, 64, 4, 4) * 0.05
    x = torch.randn(5, 64, 16, 16)
    check_conv_layer(plan.conv_s2_plan("c3", 0, 64, 128, 16), w, x, 2, 1, False)


@pytest.mark.parametrize("n", [7, 130])
def test_conv4_k4s1p0(n):
    torch.manual_seed(3)
    w = torch.randn(256, 128, 4, 4) * 0.03
    x = torch.randn(n, 128, 8, 8)
    check_conv_layer(plan.conv_k4s1p0_plan("c4", 0, 128, 256, 8), w, x, 1, 0, False)


@pytest.mark.parametrize("n", [3, 320])
def test_deconv3_pixel_pair_gradients(n):
    """deconv3's data / weight gradients over 128-byte pixel pairs of the x-padded gradient (plan.deconv_s2_plan
    extra["pair_*"], mmdyn_bn_bwd_apply_padded's layout); n = 320 selects the tile-pair kernel"""
    ops = _ops()
    torch.manual_seed(43)
    Cin, Cout, H = 64, 32, 16
    w = torch.randn(Cin, Cout, 4, 4) * 0.05
    x = torch.randn(n, Cin, H, H)
    lp = plan.deconv_s2_plan("d3", 0, Cin, Cout, H)
    xr, wr, y, dy = conv_case(lp, w, x, 2, 1, True, n)
    gpad = F.pad(dy.float(), (1, 1))                                   # one zero pixel on each side of every image row
    dA = nhwc16(gpad)
    assert tuple(dA.shape) == (n, 2 * H, 2 * H + 2, Cout)
    dx = run_fwd(lp.extra["pair_dgrad"], lp.idx_dgrad, w, dA, n)
    assert rel_err(dx.permute(0, 3, 1, 2), xr.grad) < 1e-3
    wg = lp.extra["pair_wgrad"]
    dWp = torch.zeros(wg.Cn, wg.K, dtype=torch.float32, device=DEV)
    ops.wgrad(wg, dA, nhwc16(x), dWp, n, scale=0.5, row_splits=plan.choose_row_splits(wg, n))
    dW = torch.zeros(w.numel(), dtype=torch.float32, device=DEV)
    ops.unpack_add_f32(dWp, torch.from_numpy(lp.idx_wgrad).to(DEV), dW)
    torch.cuda.synchronize()
    assert rel_err(dW, 0.5 * wr.grad.reshape(-1)) < 2e-5


def test_bn_bwd_apply_padded_matches_in_place():
    ops = _ops()
    torch.manual_seed(44)
    G, n, C, W = 2, 3, 32, 32
    rows = n * W * W
    x = torch.randn(G * rows, C, device=DEV).half()
    dy = torch.randn(G * rows, C, device=DEV).half()
    ab = torch.randn(G, C, 2, device=DEV)
    mi = torch.rand(G, C, 2, device=DEV) + 0.5
    sums2 = torch.randn(G, C, 2, device=DEV) * 10
    coef = torch.zeros(G, C, 4, device=DEV)
    dg1, db1, dg2, db2 = (torch.zeros(C, device=DEV) for _ in range(4))
    ref = dy.clone()
    ops.bn_bwd_apply(x, ab, mi, sums2, ref, dg1, db1, coef, G, rows, C, 0.25)
    out = torch.full((G * n * W, W + 2, C), 7.0, dtype=torch.float16, device=DEV)
    ops.bn_bwd_apply_padded(x, ab, mi, sums2, dy, out, 5, dg2, db2, G, rows, C, 0.25)
    torch.cuda.synchronize()
    assert torch.equal(out[:, 1:-1].reshape(-1, C), ref)               # bit-identical values, shifted by one pixel
    assert (out[:, 0] == 7.0).all() and (out[:, -1] == 7.0).all()      # the border is never written
    assert torch.equal(dg1, dg2) and torch.equal(db1, db2)


def test_pair_kernel_sizes():
    """Enough tiles for igemm_pair_kernel (two 128-row tiles per weight stage, csrc/igemm.cu): the 5x5 -> 8x8 deconv
    forward (N = 128) and data gradient (N = 256) with pixel-major tile pairs, and a Cin = 32 stride-2 conv (N = 64,
    two 64B-swizzled tap boxes per k-block) with box tile pairs."""
    torch.manual_seed(41)
    w = torch.randn(256, 128, 4, 4) * 0.03
    x = torch.randn(3072, 256, 5, 5)
    check_conv_layer(plan.deconv_k4s1p0_plan("d1", 0, 256, 128, 5), w, x, 1, 0, True)
    w = torch.randn(64, 32, 4, 4) * 0.05
    x = torch.randn(320, 32, 32, 32)
    check_conv_layer(plan.conv_s2_plan("c2", 0, 32, 64, 32), w, x, 2, 1, False)


@pytest.mark.parametrize("n", [7, 130])
def test_deconv1_k4s1p0(n):
    torch.manual_seed(4)
    w = torch.randn(256, 128, 4, 4) * 0.03
    x = torch.randn(n, 256, 5, 5)
    check_conv_layer(plan.deconv_k4s1p0_plan("d1", 0, 256, 128, 5), w, x, 1, 0, True)


def test_deconv2_s2():
    torch.manual_seed(5)
    w = torch.randn(128, 64, 4, 4) * 0.05
    x = torch.randn(5, 128, 8, 8)
    check_conv_layer(plan.deconv_s2_plan("d2", 0, 128, 64, 8), w, x, 2, 1, True)


def test_deconv3_s2():
    torch.manual_seed(6)
    w = torch.randn(64, 32, 4, 4) * 0.05
    x = torch.randn(3, 64, 16, 16)
    check_conv_layer(plan.deconv_s2_plan("d3", 0, 64, 32, 16), w, x, 2, 1, True)


def test_deconv4_out():
    torch.manual_seed(7)
    w = torch.randn(32, 3, 4, 4) * 0.1
    x = torch.randn(3, 32, 32, 32)
    check_conv_layer(plan.deconv_out_plan("d4", 0, 32, 3, 32), w, x, 2, 1, True, grad_pad=plan.LOGIT_CP, border=1)


@pytest.mark.parametrize("use_mask", [False, True])
def test_deconv4_fused_bce_epilogue(use_mask):
    """igemm out_mode 5: logits layer + BCE-with-logits (sum) + logit gradient in one launch, against
    torch conv_transpose2d + binary_cross_entropy_with_logits on the same fp16-rounded operands;
    3 groups of 2 images share the 2 target images, group 1 carries no loss, logits stored for rows 2..3."""
    ops = _ops()
    torch.manual_seed(21)
    lp = plan.deconv_out_plan("d4", 0, 32, 3, 32)
    w = torch.randn(32, 3, 4, 4) * 0.1
    x = torch.randn(6, 32, 32, 32)
    t = torch.rand(2, 3, 64, 64)
    m = (torch.rand(2, 3, 64, 64) > 0.4).float() if use_mask else None
    xr = r16(x).requires_grad_(True)
    y = F.conv_transpose2d(xr, r16(w), stride=2, padding=1)
    loss_ref = []
    for g in range(3):
        yg, tt = y[2 * g:2 * g + 2], t.double()
        if use_mask:
            yg, tt = yg * m.double(), tt * m.double()
        loss_ref.append(F.binary_cross_entropy_with_logits(yg, tt, reduction="sum"))
    (loss_ref[0] + loss_ref[2]).backward(retain_graph=True)
    dy_ref = torch.autograd.grad(loss_ref[0] + loss_ref[2], y, retain_graph=True)[0]
    logits = torch.full((6, 3, 64, 64), -77.0, device=DEV)
    dl = torch.full((6, 66, 66, plan.LOGIT_CP), 7.0, dtype=torch.float16, device=DEV)
    loss = torch.zeros(4, device=DEV)
    bce = dict(target=t.to(DEV), mask=m.to(DEV) if use_mask else None, dlogits=dl, loss=loss, gscale=2.0,
               rows_per_group=2, slots=[0, -1, 2], logit_rows=(2, 4))
    ops.igemm(lp.fwd, nhwc16(x), packed(w, lp.idx_fwd), logits, 6, bce=bce)
    torch.cuda.synchronize()
    assert rel_err(logits[2:4], y[2:4].detach()) < 5e-4
    assert (logits[:2] == -77.0).all() and (logits[4:] == -77.0).all()
    lc = loss.cpu().double()
    assert abs(lc[0] - loss_ref[0].item()) / loss_ref[0].item() < 1e-5 and lc[1] == 0 and lc[3] == 0
    assert abs(lc[2] - loss_ref[2].item()) / loss_ref[2].item() < 1e-5
    inner = dl[:, 1:65, 1:65]
    for g in (0, 2):
        assert rel_err(inner[2 * g:2 * g + 2, :, :, :3].permute(0, 3, 1, 2), 2.0 * dy_ref[2 * g:2 * g + 2]) < 1.5e-3
        assert inner[2 * g:2 * g + 2, :, :, 3:].abs().max().item() == 0
    assert (dl[2:4] == 7.0).all()                                   # group without a loss: untouched
    bord = dl.clone()
    bord[:, 1:65, 1:65] = 7.0
    assert (bord == 7.0).all()                                      # border never written


@pytest.mark.parametrize("M", [5, 128, 300])
def test_linear_fc_splitk_and_heads(M):
    ops = _ops()
    torch.manual_seed(8)
    K, N = 6400, 512
    w = torch.randn(N, K) * 0.01
    b = torch.randn(N)
    x = torch.randn(M, K)
    flat = torch.cat([w.reshape(-1), b])
    lp = plan.linear_plan("fc", [0], [N * K], K, [N])
    A = x.half().to(DEV)
    ref = F.linear(r16(x), r16(w), b.double())
    for ks in (1, plan.choose_ksplit(lp.fwd, M)):
        out = torch.zeros(M, N, dtype=torch.float32, device=DEV)
        bias = flat.to(DEV)[torch.from_numpy(lp.bias_idx).to(DEV).long()].contiguous()
        ops.igemm(lp.fwd, A, packed(flat, lp.idx_fwd), out, M, bias=bias, ksplit=ks, out_mode=2 if ks > 1 else 1)
        torch.cuda.synchronize()
        assert rel_err(out, ref) < 2e-5, (ks, rel_err(out, ref))
    # dgrad: dx = dy W
    dy = torch.randn(M, N)
    dx = torch.zeros(M, K, dtype=torch.float16, device=DEV)
    ops.igemm(lp.dgrad, dy.half().to(DEV), packed(flat, lp.idx_dgrad), dx, M)
    torch.cuda.synchronize()
    assert rel_err(dx, r16(dy) @ r16(w)) < 1e-3
    # wgrad: dW = dy^T x
    dW = torch.zeros(N, K, dtype=torch.float32, device=DEV)
    ops.wgrad(lp.wgrad, A, dy.half().to(DEV), dW, M, scale=1.0, row_splits=plan.choose_row_splits(lp.wgrad, M))
    torch.cuda.synchronize()
    assert rel_err(dW, r16(dy).t() @ r16(x)) < 2e-5


def test_upsample_linear_permuted_out():
    ops = _ops()
    torch.manual_seed(9)
    K, N, M = 256, 6400, 200
    w = torch.randn(N, K) * 0.05
    b = torch.randn(N)
    flat = torch.cat([w.reshape(-1), b])
    perm = plan.nhwc_perm(256, 5, 5)
    lp = plan.linear_plan("up", [0], [N * K], K, [N], n_perm=perm)
    z = torch.randn(M, K)
    out = torch.zeros(M, N, dtype=torch.float16, device=DEV)
    bias = flat.to(DEV)[torch.from_numpy(lp.bias_idx).to(DEV).long()].contiguous()
    ops.igemm(lp.fwd, z.half().to(DEV), packed(flat, lp.idx_fwd), out, M, bias=bias)
    torch.cuda.synchronize()
    ref = F.linear(r16(z), r16(w), b.double()).reshape(M, 256, 5, 5).permute(0, 2, 3, 1).reshape(M, -1)
    assert rel_err(out, ref) < 1e-3


def test_conv1_fwd_and_wgrad():
    ops = _ops()
    torch.manual_seed(10)
    n = 5
    w = torch.randn(32, 3, 4, 4) * 0.2
    x = torch.rand(n, 3, 64, 64)
    lp = plan.conv1_plan("c1", 0)
    out = torch.zeros(n, 32, 32, 32, dtype=torch.float16, device=DEV)
    act = torch.zeros_like(out)
    ops.conv1_fwd(x.to(DEV), packed(w, lp.idx_fwd), out, n, act=act)
    torch.cuda.synchronize()
    xr = r16(x)
    wr = r16(w).requires_grad_(True)
    y = F.conv2d(xr, wr, stride=2, padding=1)
    assert rel_err(out.permute(0, 3, 1, 2), y) < 1e-3
    # fused activation = Swish of the fp16 output, bit-identical to the stand-alone pass over it (vae.py:199)
    act2 = torch.zeros_like(out)
    ops.bn_swish_fwd(out, None, act2, 1, n * 1024, 32)
    torch.cuda.synchronize()
    assert torch.equal(act, act2)
    o64 = out.double()
    assert rel_err(act, o64 * torch.sigmoid(o64)) < 1e-3
    dy = torch.randn_like(y).half().double()
    y.backward(dy)
    dW = torch.zeros(32 * 48, dtype=torch.float32, device=DEV)
    ops.conv1_wgrad(x.to(DEV), nhwc16(dy.float()), dW, n, 0.25, 64)
    torch.cuda.synchronize()
    # the wgrad kernel reads the fp32 input directly (no fp16 rounding of x)
    wr2 = w.double().requires_grad_(True)
    F.conv2d(x.double(), wr2, stride=2, padding=1).backward(dy)
    assert rel_err(dW, 0.25 * wr2.grad.reshape(-1)) < 2e-5


@pytest.mark.parametrize("C,rows,G", [(64, 16 * 16 * 6, 1), (32, 32 * 32 * 3, 4), (256, 25 * 5, 3), (128, 64 * 9, 2)])
def test_grouped_bn_swish_fwd_bwd(C, rows, G):
    ops = _ops()
    torch.manual_seed(11)
    x = (torch.randn(G, rows, C) * 2 + 0.5).half()
    gamma, beta = torch.rand(C) + 0.5, torch.randn(C) * 0.1
    rm, rv = torch.zeros(C), torch.ones(C)
    dy = torch.randn(G, rows, C).half()
    # reference: per-group training-mode BN + swish in fp64
    xr = x.double().requires_grad_(True)
    gr, br = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm_ref, rv_ref = rm.double().clone(), rv.double().clone()
    ys = []
    for g in range(G):
        xg = xr[g].t().reshape(1, C, rows)
        yg = F.batch_norm(xg, rm_ref, rv_ref, gr, br, training=True, momentum=0.1, eps=1e-5)
        ys.append((yg * torch.sigmoid(yg)).reshape(C, rows).t())
    y = torch.stack(ys)
    y.backward(dy.double())
    # device
    xd = x.to(DEV)
    sums = torch.zeros(G, C, 2, device=DEV)
    ab = torch.empty(G, C, 2, device=DEV)
    mi = torch.empty(G, C, 2, device=DEV)
    rmd, rvd = rm.to(DEV), rv.to(DEV)
    ops.bn_stats(xd, sums, G, rows, C)
    nbt = torch.full((), 5, dtype=torch.int64, device=DEV)
    ops.bn_finalize(sums, gamma.to(DEV), beta.to(DEV), ab, mi, rmd, rvd, G, rows, C, 1e-5, 0.1, 1, nbt)
    yd = torch.empty_like(xd)
    ops.bn_swish_fwd(xd, ab, yd, G, rows, C)
    torch.cuda.synchronize()
    assert nbt.item() == 5 + G  # num_batches_tracked: one BatchNorm invocation per group
    assert rel_err(yd, y) < 1e-3
    # the single-launch form (finalize folded into the streaming kernel) must agree bit for bit
    ab2, mi2, yd2 = torch.empty_like(ab), torch.empty_like(mi), torch.empty_like(yd)
    rm2, rv2, nbt2 = rm.to(DEV), rv.to(DEV), torch.full((), 5, dtype=torch.int64, device=DEV)
    ops.bn_finalize_swish_fwd(xd, sums, gamma.to(DEV), beta.to(DEV), ab2, mi2, rm2, rv2, nbt2, yd2, G, rows, C,
                              1e-5, 0.1, 1)
    torch.cuda.synchronize()
    assert torch.equal(ab2, ab) and torch.equal(mi2, mi) and torch.equal(yd2, yd)
    assert torch.equal(rm2, rmd) and torch.equal(rv2, rvd) and nbt2.item() == 5 + G
    assert rel_err(rmd, rm_ref) < 1e-5 and rel_err(rvd, rv_ref) < 1e-5
    dyd = dy.to(DEV).clone()
    sums2 = torch.zeros(G, C, 2, device=DEV)
    dgam, dbet = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ops.bn_swish_bwd_reduce(xd, ab, mi, dyd, sums2, G, rows, C)
    coef = torch.empty(G, C, 4, device=DEV)
    ops.bn_bwd_apply(xd, ab, mi, sums2, dyd, dgam, dbet, coef, G, rows, C, 0.5)
    torch.cuda.synchronize()
    assert rel_err(dyd, xr.grad) < 3e-3
    assert rel_err(dgam, 0.5 * gr.grad) < 2e-3 and rel_err(dbet, 0.5 * br.grad) < 2e-3


def test_plain_swish_fwd_bwd():
    ops = _ops()
    torch.manual_seed(12)
    x = torch.randn(3, 100, 64).half()
    dy = torch.randn(3, 100, 64).half()
    xr = x.double().requires_grad_(True)
    y = xr * torch.sigmoid(xr)
    y.backward(dy.double())
    xd, dyd = x.to(DEV), dy.to(DEV).clone()
    yd = torch.empty_like(xd)
    ops.bn_swish_fwd(xd, None, yd, 3, 100, 64)
    ops.bn_swish_bwd_reduce(xd, None, None, dyd, None, 3, 100, 64)
    torch.cuda.synchronize()
    assert rel_err(yd, y) < 1e-3 and rel_err(dyd, xr.grad) < 1e-3


def test_swish_dropout_fwd_bwd():
    ops = _ops()
    torch.manual_seed(13)
    B, C = 37, 512
    raw = torch.randn(B, C)
    masks = [torch.empty(B, C).bernoulli_(0.9) / 0.9 for _ in range(3)] + [None]
    rr = raw.double().requires_grad_(True)
    s = rr * torch.sigmoid(rr)
    hs = torch.stack([s * (m.double() if m is not None else 1.0) for m in masks])
    dH = torch.randn(4, B, C)
    hs.backward(dH.double())
    md = [m.to(DEV) if m is not None else None for m in masks]
    h = torch.empty(4, B, C, dtype=torch.float16, device=DEV)
    ops.swish_dropout_fwd(raw.to(DEV), md, h, B, C)
    dRaw = torch.empty(B, C, dtype=torch.float16, device=DEV)
    ops.swish_dropout_bwd(raw.to(DEV), md, dH.to(DEV), dRaw, B, C)
    torch.cuda.synchronize()
    assert rel_err(h, hs) < 1e-3 and rel_err(dRaw, rr.grad) < 1e-3


def _poe_ref(mus, lvs, use_prior, eps_n):
    """vae.py:311-318 + :52-61 + problems.py:429 in fp64."""
    if not use_prior and len(mus) == 1:
        mu, lv = mus[0], lvs[0]
    else:
        mu_s = torch.stack(([torch.zeros_like(mus[0])] if use_prior else []) + list(mus))
        lv_s = torch.stack(([torch.zeros_like(lvs[0])] if use_prior else []) + list(lvs))
        e = 1e-8
        var = torch.exp(lv_s) + e
        T = 1.0 / (var + e)
        mu = torch.sum(mu_s * T, 0) / torch.sum(T, 0)
        pvar = 1.0 / torch.sum(T, 0)
        lv = torch.log(pvar + e)
    z = eps_n * torch.exp(0.5 * lv) + mu
    kl = -0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp())
    return mu, lv, z, kl


@pytest.mark.parametrize("n_exp,use_prior", [(1, False), (1, True), (2, True), (3, True)])
def test_poe_reparam_kl_fwd_bwd(n_exp, use_prior):
    ops = _ops()
    torch.manual_seed(14)
    B, D = 33, 256
    heads = [torch.randn(B, 2 * D) * 0.7 for _ in range(n_exp)]  # [mu | logvar], ld = 512
    eps_n = torch.randn(B, D)
    hr = [h.double().requires_grad_(True) for h in heads]
    mu, lv, z, kl = _poe_ref([h[:, :D] for h in hr], [h[:, D:] for h in hr], use_prior, eps_n.double())
    dz = [torch.randn(B, D), None, torch.randn(B, D)]
    klw = 0.37
    (klw * kl + (z * (dz[0] + dz[2]).double()).sum()).backward()
    hd = [h.to(DEV) for h in heads]
    mu_d, lv_d, z_d = (torch.empty(B, D, device=DEV) for _ in range(3))
    zh = torch.empty(B, D, dtype=torch.float16, device=DEV)
    kls = torch.zeros(1, device=DEV)
    ops.poe_fwd([h[:, :D] for h in hd], [h[:, D:] for h in hd], use_prior, 2 * D, eps_n.to(DEV), mu_d, lv_d, z_d, zh,
                None, kls, B, D)
    gd = [torch.zeros(B, 2 * D, device=DEV) for _ in range(n_exp)]
    ops.poe_bwd([h[:, :D] for h in hd], [h[:, D:] for h in hd], use_prior, 2 * D, eps_n.to(DEV),
                [d.to(DEV) if d is not None else None for d in dz], klw, [g[:, :D] for g in gd],
                [g[:, D:] for g in gd], 2 * D, False, B, D)
    torch.cuda.synchronize()
    assert rel_err(mu_d, mu) < 1e-5 and rel_err(lv_d, lv) < 1e-5 and rel_err(z_d, z) < 1e-5
    assert abs(kls.item() - kl.item()) / abs(kl.item()) < 1e-5
    assert rel_err(zh, z) < 1e-3
    for g, h in zip(gd, hr):
        assert rel_err(g, h.grad) < 1e-5


@pytest.mark.parametrize("use_mask", [False, True])
def test_bce_logits_and_mse(use_mask):
    ops = _ops()
    torch.manual_seed(15)
    n, HW = 5, 64 * 64
    x = torch.randn(n, 3, 64, 64) * 3
    t = torch.rand(n, 3, 64, 64)
    m = (torch.rand(n, 3, 64, 64) > 0.5).float() if use_mask else None
    xr = x.double().requires_grad_(True)
    if use_mask:
        loss = F.binary_cross_entropy_with_logits(xr * m.double(), t.double() * m.double(), reduction="sum")
    else:
        loss = F.binary_cross_entropy_with_logits(xr, t.double(), reduction="sum")
    loss.backward()
    for pad in (0, 1):
        ls = torch.zeros(1, device=DEV)
        dl = torch.full((n, 64 + 2 * pad, 64 + 2 * pad, plan.LOGIT_CP), 7.0, dtype=torch.float16, device=DEV)
        ops.bce_logits(x.to(DEV), t.to(DEV), m.to(DEV) if use_mask else None, ls, dl, 2.0, n, 64, 64, pad)
        torch.cuda.synchronize()
        assert abs(ls.item() - loss.item()) / loss.item() < 1e-5
        inner = dl[:, pad:pad + 64, pad:pad + 64]
        assert rel_err(inner[..., :3].permute(0, 3, 1, 2), 2.0 * xr.grad) < 1e-3
        assert inner[..., 3:].abs().max().item() == 0
        if pad:  # the border is never written
            bord = dl.clone()
            bord[:, pad:pad + 64, pad:pad + 64] = 7.0
            assert (bord == 7.0).all()
        # the same gradient through the stand-alone packer (loss computed outside the library)
        dl2 = torch.full_like(dl, 7.0)
        ops.logit_grad_pack(xr.grad.float().to(DEV), dl2, 2.0, n, 64, 64, pad, cp=plan.LOGIT_CP)
        torch.cuda.synchronize()
        assert rel_err(dl2[:, pad:pad + 64, pad:pad + 64, :3].permute(0, 3, 1, 2), 2.0 * xr.grad) < 1e-3
        assert (dl2[:, :pad] == 7.0).all() and (dl2[:, :, :pad] == 7.0).all()
        dl3 = torch.full((n, 64 + 2 * pad, 64 + 2 * pad, 8), 7.0, dtype=torch.float16, device=DEV)  # 8-channel form
        ops.logit_grad_pack(xr.grad.float().to(DEV), dl3, 2.0, n, 64, 64, pad, cp=8)
        torch.cuda.synchronize()
        assert rel_err(dl3[:, pad:pad + 64, pad:pad + 64, :3].permute(0, 3, 1, 2), 2.0 * xr.grad) < 1e-3
        assert dl3[:, pad:pad + 64, pad:pad + 64, 3:].abs().max().item() == 0
    # pose MSE * multiplier
    r, tt = torch.randn(n, 7), torch.rand(n, 7)
    rr = r.double().requires_grad_(True)
    l2 = 1000 * F.mse_loss(rr, tt.double(), reduction="sum")
    l2.backward()
    ls2 = torch.zeros(1, device=DEV)
    dr = torch.empty(n, 7, device=DEV)
    ops.mse(r.to(DEV), tt.to(DEV), ls2, dr, 1000.0, 0.5, n * 7)
    torch.cuda.synchronize()
    assert abs(ls2.item() - l2.item()) / l2.item() < 1e-5 and rel_err(dr, 0.5 * rr.grad) < 1e-5


@pytest.mark.parametrize("M,N,K,act", [(50, 512, 7, 1), (130, 512, 512, 0), (64, 7, 512, 0), (33, 512, 256, 1),
                                        (4100, 512, 512, 1), (2051, 256, 513, 0), (1024, 7, 512, 0), (4096, 512, 7, 1)])
def test_linear_f32(M, N, K, act):
    ops = _ops()
    torch.manual_seed(16)
    x, w, b = torch.randn(M, K), torch.randn(N, K) * 0.05, torch.randn(N)
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    y = F.linear(xr, wr, br)
    if act:
        y = F.relu(y)
    dy = torch.randn(M, N)
    y.backward(dy.double())
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    yd = torch.empty(M, N, device=DEV)
    ops.linear_f32_fwd(xd, wd, bd, yd, M, N, K, K, N, act)
    dya, dx = torch.empty(M, N, device=DEV), torch.empty(M, K, device=DEV)
    dW, db = torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
    ops.linear_f32_bwd(xd, wd, yd, dy.to(DEV), dya, dx, dW, db, M, N, K, K, N, K, act, False, 0.5)
    torch.cuda.synchronize()
    assert rel_err(yd, y) < 1e-5 and rel_err(dx, xr.grad) < 1e-5
    assert rel_err(dW, 0.5 * wr.grad) < 1e-5 and rel_err(db, 0.5 * br.grad) < 1e-5


def test_colsum_pack_gather():
    ops = _ops()
    torch.manual_seed(17)
    x = torch.randn(300, 6400).half()
    out = torch.zeros(6400, device=DEV)
    ops.colsum_f16(x.to(DEV), out, 300, 6400, 6400, 0.5)
    xf = torch.randn(77, 512)
    out2 = torch.zeros(512, device=DEV)
    ops.colsum_f32(xf.to(DEV), out2, 77, 512, 512, 2.0)
    src = torch.randn(1000)
    idx = torch.randint(-1, 1000, (333,), dtype=torch.int32)
    g = torch.empty(333, device=DEV)
    ops.gather_f32(src.to(DEV), idx.to(DEV), g)
    torch.cuda.synchronize()
    assert rel_err(out, 0.5 * x.double().sum(0)) < 1e-5
    assert rel_err(out2, 2.0 * xf.double().sum(0)) < 1e-5
    ref = torch.where(idx >= 0, src[idx.clamp(min=0).long()], torch.zeros(()))
    assert torch.equal(g.cpu(), ref)


def test_adam_and_sgd_flat_match_torch():
    ops = _ops()
    torch.manual_seed(18)
    n = 100003
    p0 = torch.randn(n)
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=1e-3)
    pd = p0.to(DEV).clone()
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        g = torch.randn(n) * (10.0 ** (step - 2))
        pt.grad = g.clone()
        opt.step()
        ops.adam_flat(pd, g.to(DEV), m, v, n, 1e-3, 0.9, 0.999, 1e-8, 0.0, step)
    torch.cuda.synchronize()
    assert (pd.cpu() - pt.detach()).abs().max().item() < 2e-6
    ps = p0.clone().requires_grad_(True)
    sgd = torch.optim.SGD([ps], lr=1e-3, momentum=0.9, weight_decay=5e-4)
    pd2, buf = p0.to(DEV).clone(), torch.zeros(n, device=DEV)
    for step in range(3):
        g = torch.randn(n)
        ps.grad = g.clone()
        sgd.step()
        ops.sgd_flat(pd2, g.to(DEV), buf, n, 1e-3, 0.9, 5e-4, step == 0)
    torch.cuda.synchronize()
    assert (pd2.cpu() - ps.detach()).abs().max().item() < 2e-6


def test_philox_rng_statistics():
    ops = _ops()
    n = 1 << 20
    a = torch.empty(n, device=DEV)
    ops.fill_normal(a, n, 1234, 0)
    b = torch.empty(n, device=DEV)
    ops.fill_normal(b, n, 1234, 0)
    c = torch.empty(n, device=DEV)
    ops.fill_normal(c, n, 1234, n)
    mk = torch.empty(n, device=DEV)
    ops.fill_dropout_mask(mk, n, 0.1, 99, 0)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1) < 5e-3
    assert abs(a.pow(4).mean().item() - 3) < 0.1
    keep = (mk > 0).float().mean().item()
    assert abs(keep - 0.9) < 2e-3
    assert torch.all((mk == 0) | ((mk - 1 / 0.9).abs() < 1e-6))


@pytest.mark.parametrize("B", [8, 100, 1024])
def test_pose_expert_on_tensor_cores_matches_fp64(B):
    """Pose MLP encoder / decoder (vae.py:118-123, 219-222, 282-283) through engine.PoseExec: the 256/512-wide Linear
    layers run on the fp16 tensor cores with the two-term split (mmdyn_split_f16 + mmdyn_igemm / mmdyn_wgrad), the
    7-wide edge layers on the fp32 SIMT kernels.  north_star tolerance for the fp32 parts: 1e-5."""
    from mmdyn_b200 import engine
    from mmdyn_b200.pytorch.models.models import setup_model
    torch.manual_seed(5)
    model = setup_model("cnn-mvae", cross_modal=True, condition_dim=0, input_dim=4096, architecture="cnn",
                        conditional=False, categorical_conditions=False, latent_size=256, use_pose=True)
    sd = {k: v.clone().double() for k, v in model.state_dict().items() if k.startswith("pose_")}
    model.to(DEV)
    arena, ex = engine.get_execs(model, torch.device(DEV))
    pex = ex["pose"]
    assert pex.tc, "the tensor-core pose path is the default"
    alloc = engine.Workspace(torch.device(DEV))
    arena.attach_grads()
    arena.grad.zero_()
    # (seed chosen so that no first-layer pre-activation lands within fp32 round-off of zero: with seed B = 1024 one of
    #  the 524,288 units sits at +1.2e-8 in fp64 and at 0 in fp32, and that single ReLU flip alone moves dW of the
    #  7 -> 512 layer by 1e-2 in the max-norm used here — tools/pose_debug.py)
    g = torch.Generator().manual_seed(B + 7)
    pose, z = torch.rand(B, 7, generator=g), torch.randn(B, 256, generator=g)
    d_heads, d_rec = torch.randn(B, 512, generator=g), 100.0 * torch.randn(B, 7, generator=g)
    # ---- fp64 reference ----
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    h1 = F.relu(F.linear(pose.double(), P["pose_encoder.fc_net.0.weight"], P["pose_encoder.fc_net.0.bias"]))
    h2 = F.linear(h1, P["pose_encoder.fc_net.2.weight"], P["pose_encoder.fc_net.2.bias"])
    heads = torch.cat([F.linear(h2, P["pose_encoder.linear_means.weight"], P["pose_encoder.linear_means.bias"]),
                       F.linear(h2, P["pose_encoder.linear_log_var.weight"], P["pose_encoder.linear_log_var.bias"])], 1)
    zr = z.double().requires_grad_(True)
    a1 = F.relu(F.linear(zr, P["pose_decoder.deconv_net.0.weight"], P["pose_decoder.deconv_net.0.bias"]))
    a2 = F.relu(F.linear(a1, P["pose_decoder.deconv_net.2.weight"], P["pose_decoder.deconv_net.2.bias"]))
    rec = F.linear(a2, P["pose_decoder.deconv_net.4.weight"], P["pose_decoder.deconv_net.4.bias"])
    for t_ in (h1, h2, a1, a2):
        t_.retain_grad()
    ((heads * d_heads.double()).sum() + (rec * d_rec.double()).sum()).backward()
    # ---- B200 path ----
    r_e = pex.enc_forward(pose.to(DEV), alloc, "penc")
    r_d = pex.dec_forward(z.to(DEV), alloc, "pdec")
    pex.enc_backward(r_e, d_heads.to(DEV), alloc, "penc", 0.5)
    dz = pex.dec_backward(r_d, d_rec.to(DEV), alloc, "pdec", 0.5)
    torch.cuda.synchronize()
    errs = {"heads": rel_err(r_e["heads"], heads), "rec": rel_err(r_d["rec"], rec), "dz": rel_err(dz, zr.grad),
            "dh2": rel_err(alloc.bufs["penc.dh2"], h2.grad), "dh1": rel_err(alloc.bufs["penc.dh1"], h1.grad),
            "da2": rel_err(alloc.bufs["pdec.da2"], a2.grad), "da1": rel_err(alloc.bufs["pdec.da1"], a1.grad)}
    named = dict(model.named_parameters())
    for k in P:
        errs[k] = rel_err(named[k].grad, 0.5 * P[k].grad)
    print(f"pose expert on tensor cores, B={B}: " + ", ".join(f"{k.replace('pose_', '')} {v:.1e}" for k, v in errs.items()))
    for k, v in errs.items():
        assert v < 1e-5, (k, v)


@pytest.mark.parametrize("M", [1000, 1024, 2048])
def test_linear_f32_first_pose_layer_without_dx(M):
    """Linear(7, 512) + ReLU backward exactly as the pose encoder calls it (dx = None)."""
    ops = _ops()
    torch.manual_seed(3)
    N, K = 512, 7
    x, w, b = torch.rand(M, K), torch.randn(N, K) * 0.3, torch.randn(N) * 0.1
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    y = F.relu(F.linear(xr, wr, br))
    dy = torch.randn(M, N)
    y.backward(dy.double())
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    yd = torch.empty(M, N, device=DEV)
    ops.linear_f32_fwd(xd, wd, bd, yd, M, N, K, K, N, 1)
    dya = torch.empty(M, N, device=DEV)
    dW, db = torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
    ops.linear_f32_bwd(xd, wd, yd, dy.to(DEV), dya, None, dW, db, M, N, K, K, N, K, 1, False, 0.5)
    torch.cuda.synchronize()
    e = (rel_err(yd, y), rel_err(dW, 0.5 * wr.grad), rel_err(db, 0.5 * br.grad))
    print(f"linear_f32 (M={M}, 512 <- 7, relu, dx=None): y {e[0]:.1e} dW {e[1]:.1e} db {e[2]:.1e}")
    assert max(e) < 1e-5, e
