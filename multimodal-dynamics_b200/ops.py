"""Thin torch-tensor front end over the C ABI (lib.py): extracts device pointers / the current
CUDA stream and fills the descriptor structs.  No math happens here and nothing falls back to
PyTorch: every function launches one or two kernels of libmmdyn_b200.so or raises."""
import ctypes as C

import torch

from . import lib as _lib
from .lib import IgemmDesc, WgradDesc, MAX_GROUPS, check


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "mmdyn_b200 kernels need CUDA tensors (there is no CPU path)"
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _L():
    return _lib.init(torch.cuda.current_device())


def _ptr_array(tensors, n):
    arr = (C.c_void_p * n)()
    for i in range(n):
        arr[i] = tensors[i].data_ptr() if (i < len(tensors) and tensors[i] is not None) else None
    return arr


# ---------------------------------------------------------------------------------------------
# optional per-launch timing with CUDA events on the launching stream (bench.py roofline pass);
# off by default: the timed steps of the benchmark run without it
# ---------------------------------------------------------------------------------------------
_prof = None


def start_profile():
    global _prof
    _prof = []


def stop_profile():
    """-> {tag: dict(count, ms, flops, bytes)} (synchronises)."""
    global _prof
    rec, _prof = _prof, None
    torch.cuda.synchronize()
    out = {}
    for tag, e0, e1, alg in rec or []:
        d = out.setdefault(tag, dict(count=0, ms=0.0, flops=0.0, bytes=0.0))
        d["count"] += 1
        d["ms"] += e0.elapsed_time(e1)
        if alg is not None:
            d["flops"] += alg[0]
            d["bytes"] += alg[1]
    return out


_FAMILY = {"adam_flat": "adam_kernel", "sgd_flat": "sgd_kernel", "bn_stats": "bn_stats_bulk_kernel",
           "bn_swish_fwd": "bn_swish_fwd_bulk_kernel", "bn_swish_bwd_reduce": "bn_swish_bwd_reduce_bulk_kernel",
           "bn_bwd_apply": "bn_bwd_apply_bulk_kernel", "linear_f32_fwd": "sgemm_kernel (pose MLP forward)",
           "linear_f32_bwd": "sgemm_kernel (pose MLP backward)", "conv1_fwd": "conv1_fwd_kernel",
           "conv1.wgrad": "wgrad_kernel<32> (conv1, Cin = 3 -> 8-channel pixels)"}


_PATCH_TAGS = set()  # layer tags that run on igemm_patch_kernel (filled by igemm() as launches happen)


def kernel_family(tag):
    """Kernel NAME a profile tag runs in: the GEMM launches are tagged per layer ("deconv3.fwd") and pooled here
    into the kernel template that executes them; the streaming kernels are tagged by their own name."""
    if tag in _FAMILY:
        return _FAMILY[tag]
    if tag in _PATCH_TAGS:
        return "igemm_patch_kernel<*> (merged 3x3-tap layers: deconv2/3/4 forward, conv2/3 dgrad)"
    if tag.endswith(".fwd") or tag.endswith(".dgrad"):
        return "igemm_tma_kernel<*> + igemm_pair_kernel<*> (conv / deconv / linear forward + dgrad)"
    if tag.endswith(".wgrad"):
        return "wgrad_tma_kernel<*> (weight gradients)"
    return tag + "_kernel"


class _Timed:
    __slots__ = ("tag", "alg", "e0")

    def __init__(self, tag, alg=None):
        self.tag, self.alg = tag, alg

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if _prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.tag, self.e0, e1, self.alg() if callable(self.alg) else self.alg))


def _esz(t):
    return t.element_size()


# Branch staggering (engine.StepEngine._fork): the visual and tactile branches run the same kernel sequence, so without an
# offset their tensor-bound GEMMs (and their HBM-bound BatchNorm passes) hit the machine at the same time.  The first branch
# records an event after its k-th GEMM launch; the next branch starts behind that event.
_stagger = None  # [event, launches to go]


def arm_stagger(event, after):
    global _stagger
    _stagger = [event, int(after)] if event is not None else None


def _stagger_tick():
    global _stagger
    if _stagger is not None:
        _stagger[1] -= 1
        if _stagger[1] <= 0:
            _stagger[0].record()
            _stagger = None


def igemm(geom, A, Wp, out, n_img, bias=None, ksplit=1, out_mode=None, ldc=None, a_pix_stride=None, tag="igemm",
          macs_per_img=None, bce=None, stats=None):
    """bce (out_mode 5, the logits layer with its loss fused): dict(target, mask, dlogits, loss, gscale,
    rows_per_group, slots=[loss index per group or -1], logit_rows=(lo, hi))."""
    d = IgemmDesc()
    d.A, d.W, d.out = A.data_ptr(), Wp.data_ptr(), out.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    d.n_img, d.P, d.OXv, d.IH, d.IW = n_img, geom.P, geom.OXv, geom.IH, geom.IW
    d.a_pix_stride = geom.a_pix_stride if a_pix_stride is None else a_pix_stride
    d.Cin, d.s_in, d.ntaps, d.n_phases = geom.Cin, geom.s_in, geom.ntaps, geom.n_phases
    for p in range(geom.n_phases):
        for t in range(geom.ntaps):
            d.tap_dy[p][t] = geom.tap_dy[p][t]
            d.tap_dx[p][t] = geom.tap_dx[p][t]
        d.off_y[p], d.off_x[p] = geom.off_y[p], geom.off_x[p]
    d.N, d.block_n, d.ksplit, d.row_mode = geom.N, geom.block_n, ksplit, geom.row_mode
    d.out_mode = geom.out_mode if out_mode is None else out_mode
    d.OH, d.OW, d.s_out = geom.OH, geom.OW, geom.s_out
    d.ldc = geom.ldc if ldc is None else ldc
    d.a_row_stride, d.a_img_stride = geom.a_row_stride, geom.a_img_stride
    d.s_in_x = int(getattr(geom, "s_in_x", 0))
    d.patch_mode = int(getattr(geom, "patch", 0))
    if d.patch_mode:
        _PATCH_TAGS.add(tag)
    if stats is not None:  # (sums [G][C][2] fp32 zeroed, images per group): BatchNorm statistics in the epilogue
        assert (d.patch_mode and d.out_mode == 4) or (not d.patch_mode and d.out_mode == 0)
        d.bn_sums, d.bn_rows_per_group = stats[0].data_ptr(), int(stats[1])
    if bce is not None:
        d.out_mode = 5
        d.bce_target, d.bce_mask = bce["target"].data_ptr(), _ptr(bce.get("mask"))
        d.bce_dlogits, d.bce_loss = _ptr(bce.get("dlogits")), bce["loss"].data_ptr()
        d.bce_gscale, d.bce_rows_per_group = bce["gscale"], bce["rows_per_group"]
        for g in range(MAX_GROUPS):
            d.bce_slot[g] = bce["slots"][g] if g < len(bce["slots"]) else -1
        d.logit_row_lo, d.logit_row_hi = bce["logit_rows"]

    def alg():
        # algorithmic work: true MACs of the layer (no padding / phase-union waste) and one read of
        # each operand + one write of the output
        macs = (macs_per_img if macs_per_img is not None else geom.P * geom.n_phases * geom.N * geom.K) * n_img
        a_elems = geom.a_img_stride if geom.a_img_stride else geom.IH * geom.IW * geom.Cin
        nbytes = n_img * a_elems * 2 + Wp.numel() * 2 + out.numel() * _esz(out)
        return 2.0 * macs, float(nbytes)
    with _Timed(tag, alg):
        check(_L().mmdyn_igemm(C.byref(d), _stream()), "mmdyn_igemm")
    _stagger_tick()


def wgrad(geom, G, Nat, dW, n_img, scale=1.0, row_splits=1, ldw=None, nat_stride=None, g_pix_stride=None,
          tag="wgrad", macs_per_img=None):
    d = WgradDesc()
    d.G, d.Nat, d.dW = G.data_ptr(), Nat.data_ptr(), dW.data_ptr()
    d.n_img, d.P, d.OXv, d.IH, d.IW = n_img, geom.P, geom.OXv, geom.IH, geom.IW
    d.g_pix_stride = geom.g_pix_stride if g_pix_stride is None else g_pix_stride
    d.Cg, d.s_in, d.ntaps = geom.Cg, geom.s_in, geom.ntaps
    for t in range(geom.ntaps):
        d.tap_dy[t], d.tap_dx[t] = geom.tap_dy[t], geom.tap_dx[t]
    d.Cn = geom.Cn
    d.nat_stride = geom.nat_stride if nat_stride is None else nat_stride
    d.ldw = geom.K if ldw is None else ldw
    d.row_splits, d.scale = row_splits, scale
    d.g_row_stride, d.g_img_stride = geom.g_row_stride, geom.g_img_stride
    d.s_in_x = int(getattr(geom, "s_in_x", 0))

    def alg():
        macs = (macs_per_img if macs_per_img is not None else geom.P * geom.Cn * geom.K) * n_img
        g_elems = geom.g_img_stride if geom.g_img_stride else geom.IH * geom.IW * geom.Cg
        nbytes = n_img * g_elems * 2 + n_img * geom.P * geom.Cn * 2 + dW.numel() * 4
        return 2.0 * macs, float(nbytes)
    with _Timed(tag, alg):
        check(_L().mmdyn_wgrad(C.byref(d), _stream()), "mmdyn_wgrad")


def conv1_fwd(x, Wp, out, n_img, act=None):
    """act: optional second output, Swish(out) (the layer's activation, vae.py:199)"""
    with _Timed("conv1_fwd", lambda: (2.0 * n_img * 1024 * 32 * 48,
                                      n_img * (3 * 64 * 64 * 4 + 1024 * 32 * (4.0 if act is not None else 2.0)))):
        check(_L().mmdyn_conv1_fwd(_ptr(x), _ptr(Wp), _ptr(out), _ptr(act), n_img, _stream()), "conv1_fwd")


def conv1_wgrad(x, dRaw, dW, n_img, scale, row_splits):
    with _Timed("conv1_wgrad", lambda: (2.0 * n_img * 1024 * 32 * 48, n_img * (3 * 64 * 64 * 4 + 1024 * 32 * 2.0))):
        check(_L().mmdyn_conv1_wgrad(_ptr(x), _ptr(dRaw), _ptr(dW), n_img, scale, row_splits, _stream()), "conv1_wgrad")


def bn_stats(x, sums, G, rows, Cch):
    with _Timed("bn_stats", lambda: (0.0, G * rows * Cch * 2.0)):
        check(_L().mmdyn_bn_stats(_ptr(x), _ptr(sums), G, rows, Cch, _stream()), "bn_stats")


def bn_finalize_swish_fwd(x, sums, gamma, beta, ab, mean_invstd, running_mean, running_var, num_batches_tracked, y, G,
                          rows, Cch, eps, momentum, stat_repeat=1):
    with _Timed("bn_swish_fwd", lambda: (0.0, G * rows * Cch * 4.0)):
        check(_L().mmdyn_bn_finalize_swish_fwd(_ptr(x), _ptr(sums), _ptr(gamma), _ptr(beta), _ptr(ab), _ptr(mean_invstd),
                                               _ptr(running_mean), _ptr(running_var), _ptr(num_batches_tracked),
                                               _ptr(y), G, rows, Cch, eps, momentum, stat_repeat, _stream()),
              "bn_finalize_swish_fwd")


def bn_finalize(sums, gamma, beta, ab, mean_invstd, running_mean, running_var, G, rows, Cch, eps, momentum,
                stat_repeat=1, num_batches_tracked=None):
    with _Timed("bn_finalize", None):
        check(_L().mmdyn_bn_finalize(_ptr(sums), _ptr(gamma), _ptr(beta), _ptr(ab), _ptr(mean_invstd),
                                     _ptr(running_mean), _ptr(running_var), G, rows, Cch, eps, momentum, stat_repeat,
                                     _ptr(num_batches_tracked), _stream()),
              "bn_finalize")


def bn_swish_fwd(x, ab, y, G, rows, Cch):
    with _Timed("bn_swish_fwd", lambda: (0.0, G * rows * Cch * 4.0)):
        check(_L().mmdyn_bn_swish_fwd(_ptr(x), _ptr(ab), _ptr(y), G, rows, Cch, _stream()), "bn_swish_fwd")


def bn_swish_bwd_reduce(x, ab, mean_invstd, dY, sums2, G, rows, Cch):
    with _Timed("bn_swish_bwd_reduce", lambda: (0.0, G * rows * Cch * (4.0 if ab is not None else 6.0))):
        check(_L().mmdyn_bn_swish_bwd_reduce(_ptr(x), _ptr(ab), _ptr(mean_invstd), _ptr(dY), _ptr(sums2), G, rows,
                                             Cch, _stream()), "bn_swish_bwd_reduce")


def bn_bwd_apply(x, ab, mean_invstd, sums2, dU, dgamma, dbeta, coef, G, rows, Cch, unscale):
    with _Timed("bn_bwd_apply", lambda: (0.0, G * rows * Cch * 6.0)):
        check(_L().mmdyn_bn_bwd_apply(_ptr(x), _ptr(ab), _ptr(mean_invstd), _ptr(sums2), _ptr(dU), _ptr(dgamma),
                                      _ptr(dbeta), _ptr(coef), G, rows, Cch, unscale, _stream()), "bn_bwd_apply")


def bn_bwd_apply_padded(x, ab, mean_invstd, sums2, dU, dX_padded, row_w_log2, dgamma, dbeta, G, rows, Cch, unscale):
    """dX into [rows / 2^w][2^w + 2][C] (one zero pixel on each side of every image row) instead of over dU"""
    with _Timed("bn_bwd_apply", lambda: (0.0, G * rows * Cch * 6.0)):
        check(_L().mmdyn_bn_bwd_apply_padded(_ptr(x), _ptr(ab), _ptr(mean_invstd), _ptr(sums2), _ptr(dU), _ptr(dX_padded),
                                             row_w_log2, _ptr(dgamma), _ptr(dbeta), G, rows, Cch, unscale, _stream()),
              "bn_bwd_apply_padded")


def swish_dropout_fwd(raw, masks, h, B, Cch):
    n = len(masks)
    with _Timed("swish_dropout_fwd", None):
        check(_L().mmdyn_swish_dropout_fwd(_ptr(raw), _ptr_array(masks, n), _ptr(h), n, B, Cch, _stream()),
              "swish_dropout_fwd")


def swish_dropout_bwd(raw, masks, dH, dRaw, B, Cch):
    n = len(masks)
    with _Timed("swish_dropout_bwd", None):
        check(_L().mmdyn_swish_dropout_bwd(_ptr(raw), _ptr_array(masks, n), _ptr(dH), _ptr(dRaw), n, B, Cch,
                                           _stream()), "swish_dropout_bwd")


def poe_fwd(mu_e, lv_e, use_prior, ld, eps, mu, lv, z, zh, zh2, kl_sum, B, D):
    n = len(mu_e)
    with _Timed("poe_fwd", None):
        check(_L().mmdyn_poe_fwd(_ptr_array(mu_e, 4), _ptr_array(lv_e, 4), n, int(use_prior), ld, _ptr(eps), _ptr(mu),
                                 _ptr(lv), _ptr(z), _ptr(zh), _ptr(zh2), _ptr(kl_sum), B, D, _stream()), "poe_fwd")


def poe_bwd(mu_e, lv_e, use_prior, ld, eps, dzs, kl_coef, dmu_e, dlv_e, ld_out, accumulate, B, D,
            dmu_in=None, dlv_in=None):
    n = len(mu_e)
    with _Timed("poe_bwd", None):
        check(_L().mmdyn_poe_bwd(_ptr_array(mu_e, 4), _ptr_array(lv_e, 4), n, int(use_prior), ld, _ptr(eps),
                                 _ptr_array(dzs, 3), _ptr(dmu_in), _ptr(dlv_in), kl_coef, _ptr_array(dmu_e, 4),
                                 _ptr_array(dlv_e, 4), ld_out,
                                 int(accumulate), B, D, _stream()), "poe_bwd")


def _poe_passes(passes):
    from .lib import PoePass
    arr = (PoePass * len(passes))()
    dp = lambda t: t.data_ptr() if t is not None else None
    for k, a in enumerate(passes):
        q = arr[k]
        q.n_experts = len(a["mu_e"])
        for e in range(q.n_experts):
            q.mu_e[e], q.lv_e[e] = a["mu_e"][e].data_ptr(), a["lv_e"][e].data_ptr()
            if "dmu_e" in a:
                q.dmu_e[e], q.dlv_e[e] = a["dmu_e"][e].data_ptr(), a["dlv_e"][e].data_ptr()
        q.eps = a["eps"].data_ptr()
        for f in ("mu", "lv", "z", "zh", "zh2", "kl_sum", "dmu_in", "dlv_in"):
            setattr(q, f, dp(a.get(f)))
        for j, t in enumerate(a.get("dz", [])[:3]):
            q.dz[j] = dp(t)
    return arr


def poe_fwd_multi(passes, use_prior, ld, B, D):
    """passes: list of dict(mu_e, lv_e, eps, mu, lv, z, zh, zh2, kl_sum) — every pass of the step in one launch."""
    with _Timed("poe_fwd", None):
        check(_L().mmdyn_poe_fwd_multi(_poe_passes(passes), len(passes), int(use_prior), ld, B, D, _stream()), "poe_fwd_multi")


def poe_bwd_multi(passes, use_prior, ld, kl_coef, ld_out, accumulate, B, D):
    """passes: list of dict(mu_e, lv_e, eps, dz, dmu_e, dlv_e[, dmu_in, dlv_in]) — one launch for all passes."""
    with _Timed("poe_bwd", None):
        check(_L().mmdyn_poe_bwd_multi(_poe_passes(passes), len(passes), int(use_prior), ld, kl_coef, ld_out,
                                       int(accumulate), B, D, _stream()), "poe_bwd_multi")


def bce_logits(logits, target, mask, loss_sum, dlogits, gscale, n, H, W, pad=0):
    """pad: border (pixels) of the NHWC4 gradient images, see include/mmdyn_b200.h"""
    with _Timed("bce_logits", lambda: (0.0, n * H * W * (3 * 8.0 + (8.0 if dlogits is not None else 0.0)))):
        check(_L().mmdyn_bce_logits(_ptr(logits), _ptr(target), _ptr(mask), _ptr(loss_sum), _ptr(dlogits), gscale, n,
                                    H, W, pad, _stream()), "bce_logits")


def bce_logits_flat(logits, target, mask, loss_sum, per_sample_sum, dlogits, gscale, n, per_sample):
    with _Timed("bce_logits_flat", None):
        check(_L().mmdyn_bce_logits_flat(_ptr(logits), _ptr(target), _ptr(mask), _ptr(loss_sum), _ptr(per_sample_sum),
                                         _ptr(dlogits), gscale, n, per_sample, _stream()), "bce_logits_flat")


def mse_rows(recon, target, row_sum, mult, n, d):
    with _Timed("mse_rows", None):
        check(_L().mmdyn_mse_rows(_ptr(recon), _ptr(target), _ptr(row_sum), mult, n, d, _stream()), "mse_rows")


def mse(recon, target, loss_sum, drecon, mult, gscale, n):
    with _Timed("mse", None):
        check(_L().mmdyn_mse(_ptr(recon), _ptr(target), _ptr(loss_sum), _ptr(drecon), mult, gscale, n, _stream()), "mse")


def linear_f32_fwd(x, W, b, y, M, N, K, ldx, ldy, act):
    with _Timed("linear_f32_fwd", None):
        check(_L().mmdyn_linear_f32_fwd(_ptr(x), _ptr(W), _ptr(b), _ptr(y), M, N, K, ldx, ldy, act, _stream()),
              "linear_f32_fwd")


def linear_f32_bwd(x, W, y, dy, dy_act, dx, dW, db, M, N, K, ldx, ldy, lddx, act, dx_accumulate, scale):
    with _Timed("linear_f32_bwd", None):
        check(_L().mmdyn_linear_f32_bwd(_ptr(x), _ptr(W), _ptr(y), _ptr(dy), _ptr(dy_act), _ptr(dx), _ptr(dW), _ptr(db),
                                        M, N, K, ldx, ldy, lddx, act, int(dx_accumulate), scale, _stream()),
              "linear_f32_bwd")


def colsum_f32(x, out, M, N, ld, scale):
    with _Timed("colsum_f32", None):
        check(_L().mmdyn_colsum_f32(_ptr(x), _ptr(out), M, N, ld, scale, _stream()), "colsum_f32")


def colsum_f16(x, out, M, N, ld, scale):
    with _Timed("colsum_f16", None):
        check(_L().mmdyn_colsum_f16(_ptr(x), _ptr(out), M, N, ld, scale, _stream()), "colsum_f16")


def pack_f16(src, idx, dst):
    with _Timed("pack_f16", lambda: (0.0, idx.numel() * 10.0)):
        check(_L().mmdyn_pack_f16(_ptr(src), _ptr(idx), _ptr(dst), idx.numel(), _stream()), "pack_f16")


def gather_f32(src, idx, dst):
    with _Timed("gather_f32", None):
        check(_L().mmdyn_gather_f32(_ptr(src), _ptr(idx), _ptr(dst), idx.numel(), _stream()), "gather_f32")


def gather_add_f32(src, inv, dst):
    """dst[k] += src[inv[k]] where inv[k] >= 0 (dst: a slice of the gradient arena)."""
    with _Timed("unpack_add_f32", lambda: (0.0, inv.numel() * 16.0)):
        check(_L().mmdyn_gather_add_f32(_ptr(src), _ptr(inv), _ptr(dst), inv.numel(), _stream()), "gather_add_f32")


def unpack_add_f32(src, idx, dst):
    with _Timed("unpack_add_f32", lambda: (0.0, idx.numel() * 16.0)):
        check(_L().mmdyn_unpack_add_f32(_ptr(src), _ptr(idx), _ptr(dst), idx.numel(), _stream()), "unpack_add_f32")


def f32_to_f16(src, dst, n, scale=1.0):
    with _Timed("f32_to_f16", None):
        check(_L().mmdyn_f32_to_f16(_ptr(src), _ptr(dst), n, scale, _stream()), "f32_to_f16")


def scale_f32(x, n, s):
    with _Timed("scale_f32", None):
        check(_L().mmdyn_scale_f32(_ptr(x), n, s, _stream()), "scale_f32")


def logit_grad_pack(dl, out, scale, n, H, W, pad=0, cp=8):
    """cp: channels per pixel of `out` (3 used): 4 = the layout of the logit gradients (plan.LOGIT_CP), 8 = the repacked input
    image of the first conv's weight gradient"""
    with _Timed("logit_grad_pack", None):
        check(_L().mmdyn_logit_grad_pack(_ptr(dl), _ptr(out), scale, n, H, W, pad, cp, _stream()), "logit_grad_pack")


def adam_flat(p, g, m, v, n, lr, b1, b2, eps, wd, step, gscale=1.0):
    with _Timed("adam_flat", lambda: (0.0, n * 28.0)):
        check(_L().mmdyn_adam_flat(_ptr(p), _ptr(g), _ptr(m), _ptr(v), n, lr, b1, b2, eps, wd, step, gscale, _stream()),
              "adam_flat")


def adam_flat_devstep(p, g, m, v, n, lr, b1, b2, eps, wd, step_dev, gscale=1.0, flag=None):
    with _Timed("adam_flat", lambda: (0.0, n * 28.0)):
        check(_L().mmdyn_adam_flat_guarded(_ptr(p), _ptr(g), _ptr(m), _ptr(v), n, lr, b1, b2, eps, wd, _ptr(step_dev),
                                           gscale, _ptr(flag) if flag is not None else None, _stream()),
              "adam_flat_guarded")


def sgd_flat(p, g, buf, n, lr, momentum, wd, first_step, gscale=1.0, flag=None):
    with _Timed("sgd_flat", None):
        check(_L().mmdyn_sgd_flat_guarded(_ptr(p), _ptr(g), _ptr(buf), n, lr, momentum, wd, int(first_step), gscale,
                                          _ptr(flag) if flag is not None else None, _stream()), "sgd_flat_guarded")


def fill_normal(out, n, seed, offset, ctr=None):
    with _Timed("fill_normal", None):
        check(_L().mmdyn_fill_normal(_ptr(out), n, seed, offset, _ptr(ctr), _stream()), "fill_normal")


def fill_dropout_mask(out, n, p_drop, seed, offset, ctr=None):
    with _Timed("fill_dropout_mask", None):
        check(_L().mmdyn_fill_dropout_mask(_ptr(out), n, p_drop, seed, offset, _ptr(ctr), _stream()), "fill_dropout_mask")


def rng_advance(ctr, inc):
    with _Timed("rng_advance", None):
        check(_L().mmdyn_rng_advance(_ptr(ctr), inc, _stream()), "rng_advance")


def resize_table(in_h, in_w, out_h, out_w):
    """Host-side coefficient table of the PIL-exact bilinear resize (int32 CPU tensor; no GPU needed)."""
    L = _lib.load()  # host-only entry points: no device initialisation
    n = L.mmdyn_resize_table_ints(in_h, in_w, out_h, out_w)
    if n <= 0:
        raise ValueError(f"resize_table: bad sizes {(in_h, in_w, out_h, out_w)}")
    t = torch.empty(n, dtype=torch.int32)
    check(L.mmdyn_resize_table(in_h, in_w, out_h, out_w, t.data_ptr(), n), "resize_table")
    return t


def frames_u8_to_f32(frames, index, table_dev, out):
    """out[i] = ToTensor(Resize(frames[index[i]])): frames uint8 (N,H,W,3), index int64 (n,) or None, out fp32 (n,3,h,w)."""
    n, _, oh, ow = out.shape
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3 and frames.is_contiguous()
    assert index is None or (index.dtype == torch.int64 and index.numel() == n)
    with _Timed("frames_u8_to_f32", lambda: (0.0, n * (frames.shape[1] * frames.shape[2] * 3.0 + 3 * oh * ow * 4.0))):
        for s in range(0, n, 32768):  # grid.y limit
            e = min(n, s + 32768)
            check(_L().mmdyn_frames_u8_to_f32(_ptr(frames), _ptr(index[s:e]) if index is not None else
                                              (None if s == 0 else _ptr(torch.arange(s, e, device=out.device))),
                                              _ptr(table_dev), _ptr(out[s:e]), e - s, frames.shape[1], frames.shape[2],
                                              oh, ow, _stream()), "frames_u8_to_f32")


def linear_f32_acc(x, W, y, M, N, K, ldx, ldw, ldy):
    with _Timed("cond_linear", None):
        check(_L().mmdyn_linear_f32_acc(_ptr(x), _ptr(W), _ptr(y), M, N, K, ldx, ldw, ldy, _stream()), "linear_f32_acc")


def linear_f32_wgrad(x, dy, dW, M, N, K, ldx, lddy, ldw, scale):
    with _Timed("cond_linear", None):
        check(_L().mmdyn_linear_f32_wgrad(_ptr(x), _ptr(dy), _ptr(dW), M, N, K, ldx, lddy, ldw, scale, _stream()),
              "linear_f32_wgrad")


def cond_add_f16(raw, c, W, n_idx, R, N, ldw, col0, cd):
    with _Timed("cond_linear", None):
        check(_L().mmdyn_cond_add_f16(_ptr(raw), _ptr(c), _ptr(W), _ptr(n_idx), R, N, ldw, col0, cd, _stream()),
              "cond_add_f16")


def cond_wgrad_f16(g, c, dW, n_idx, R, N, ldw, col0, cd, scale):
    with _Timed("cond_linear", None):
        check(_L().mmdyn_cond_wgrad_f16(_ptr(g), _ptr(c), _ptr(dW), _ptr(n_idx), R, N, ldw, col0, cd, scale, _stream()),
              "cond_wgrad_f16")


def relu_f32(x, y):
    with _Timed("relu_f32", None):
        check(_L().mmdyn_relu_f32(_ptr(x), _ptr(y), x.numel(), _stream()), "relu_f32")


def act_grad_f32(y, dy, dx, M, N, ldy, act):
    with _Timed("act_grad_f32", None):
        check(_L().mmdyn_act_grad_f32(_ptr(y), _ptr(dy), _ptr(dx), M, N, ldy, act, _stream()), "act_grad_f32")


def ipc_export(t):
    """(64-byte CUDA IPC handle, byte offset) of a device tensor's memory — picklable, for the other ranks."""
    h = C.create_string_buffer(64)
    off = C.c_longlong(0)
    check(_L().mmdyn_ipc_export(_ptr(t), h, C.byref(off)), "ipc_export")
    return bytes(h.raw), int(off.value)


def ipc_import(handle, offset):
    """Raw device pointer (int) of a peer rank's exported memory, mapped for kernels of the CURRENT device."""
    p = C.c_void_p(0)
    check(_L().mmdyn_ipc_import(C.create_string_buffer(handle, 64), int(offset), C.byref(p)), "ipc_import")
    return int(p.value)


def split_f16(x, out, M, N, mode, relu=False, mask_y=None, x_out=None, colsum0=None, colsum1=None, n_split=0,
              colsum_scale=1.0):
    """fp32 [M][N] -> fp16 [M][3N] two-term split ([hi|lo|hi] mode 0, [hi|hi|lo] mode 1), see include/mmdyn_b200.h."""
    with _Timed("split_f16", None):
        check(_L().mmdyn_split_f16(_ptr(x), _ptr(mask_y), _ptr(x_out), _ptr(out), M, N, mode, int(relu), _ptr(colsum0),
                                   _ptr(colsum1), n_split, colsum_scale, _stream()), "split_f16")


def enable_peer_access(peer_device):
    check(_L().mmdyn_enable_peer_access(int(peer_device)), "enable_peer_access")


def peer_rs_adam_ag(grad_ptrs, param_ptrs, flag_ptrs, m, v, n, rank, world, lr, b1, b2, eps, wd, step_dev, epoch_dev,
                    gscale, flag, counter):
    """Fused reduce-scatter + Adam + all-gather over peer memory (csrc/peer.cu).  *_ptrs: python lists of the N ranks'
    raw device pointers (peer-mapped; own rank included).  Algorithmic bytes: own shard 28 B/param + peer traffic."""
    arr = lambda ps: (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in ps])
    with _Timed("peer_rs_adam_ag", lambda: (0.0, n * 4.0 * (2.0 * (world - 1) / world) + n * 28.0 / world)):
        check(_L().mmdyn_peer_rs_adam_ag(arr(grad_ptrs), arr(param_ptrs), arr(flag_ptrs), _ptr(m), _ptr(v), n, rank, world,
                                         lr, b1, b2, eps, wd, _ptr(step_dev), _ptr(epoch_dev), gscale, _ptr(flag),
                                         _ptr(counter), _stream()), "peer_rs_adam_ag")
