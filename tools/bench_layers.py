"""Micro-benchmark of single tensor-core layers through the C ABI (random operands, full-size row counts):
time per launch, algorithmic TFLOP/s, microseconds per 128-row tile and SM.  Used for kernel A/B work and as
the `<cmd>` of ncu captures of one kernel:

    python tools/bench_layers.py [name ...]        # e.g. deconv3.fwd deconv4.fwd; no names = all
    MMDYN_NO_PATCH=1 python tools/bench_layers.py  # generic one-box-per-tap kernel for the merged layers
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdyn_b200 import ops, plan  # noqa: E402

DEV = "cuda"


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(name, lp, which, n_img, in_shape, out_shape, out_dtype=torch.float16, stats_c=0, bce=False, geom=None):
    geom = geom or (lp.fwd if which == "fwd" else lp.dgrad)
    idx = lp.idx_fwd if which == "fwd" else lp.idx_dgrad
    A = torch.randn(n_img, *in_shape, device=DEV).half()
    W = (torch.randn(idx.shape, device=DEV) * 0.05).half()
    W[torch.from_numpy(idx).to(DEV) < 0] = 0
    out = torch.zeros(n_img, *out_shape, device=DEV, dtype=out_dtype)
    kw = {}
    if stats_c and getattr(geom, "patch", 0):
        sums = torch.zeros(4, stats_c, 2, device=DEV)
        kw["stats"] = (sums, n_img // 4)
    if bce:  # the logits layer as the step runs it: BCE loss + logit gradient fused into the epilogue, 4 groups
        B = n_img // 4
        kw["bce"] = dict(target=torch.rand(B, 3, 64, 64, device=DEV), mask=None, loss=torch.zeros(64, device=DEV), gscale=1.0,
                         dlogits=torch.zeros(n_img, 66, 66, 4, device=DEV, dtype=torch.float16), rows_per_group=B,
                         slots=[8, 9, 10, 11], logit_rows=(0, B))
    ms = timeit(lambda: ops.igemm(geom, A, W, out, n_img, **kw))
    rows = n_img * geom.P * geom.n_phases
    tiles = (rows + 127) // 128 * (geom.N // geom.block_n)
    macs = lp.extra["macs"] * n_img
    abytes = A.numel() * 2 + out.numel() * out.element_size()
    print(f"{name:16s} n={n_img:6d} ms={ms:7.3f} tiles={tiles:7d} kb/tile={geom.K // 64:3d} us/tile/SM={ms * 1e3 / tiles * 148:7.2f} "
          f"TFLOPs={2 * macs / ms / 1e9:7.1f} algGB/s={abytes / ms / 1e6:7.0f} patch={getattr(geom, 'patch', 0)}", flush=True)


def run_wgrad(name, lp, n_img, g_shape, nat_shape, wg=None):
    wg = wg or lp.wgrad
    G = torch.randn(n_img, *g_shape, device=DEV).half()
    Nat = torch.randn(n_img, *nat_shape, device=DEV).half()
    dW = torch.zeros(wg.Cn, wg.K, device=DEV)
    rs = plan.choose_row_splits(wg, n_img)
    ms = timeit(lambda: ops.wgrad(wg, G, Nat, dW, n_img, scale=1.0, row_splits=rs))
    macs = lp.extra["macs"] * n_img
    abytes = G.numel() * 2 + Nat.numel() * 2
    print(f"{name:16s} n={n_img:6d} ms={ms:7.3f} row_splits={rs:4d} K={wg.K:5d} Cn={wg.Cn:4d} TFLOPs={2 * macs / ms / 1e9:7.1f} "
          f"algGB/s={abytes / ms / 1e6:7.0f}", flush=True)


CASES = {
    "deconv4.fwd": lambda R: run("deconv4.fwd", plan.deconv_out_plan("d4", 0), "fwd", R, (32, 32, 32), (3, 64, 64), torch.float32),
    "deconv4.fwd.bce": lambda R: run("deconv4.fwd.bce", plan.deconv_out_plan("d4", 0), "fwd", R, (32, 32, 32), (3, 64, 64), torch.float32, bce=True),
    "deconv3.fwd": lambda R: run("deconv3.fwd", plan.deconv_s2_plan("d3", 0, 64, 32, 16), "fwd", R, (16, 16, 64), (32, 32, 32), stats_c=32),
    "deconv3.fwd.nostats": lambda R: run("deconv3.fwd.nost", plan.deconv_s2_plan("d3", 0, 64, 32, 16), "fwd", R, (16, 16, 64), (32, 32, 32)),
    "deconv2.fwd": lambda R: run("deconv2.fwd", plan.deconv_s2_plan("d2", 0, 128, 64, 8), "fwd", R, (8, 8, 128), (16, 16, 64), stats_c=64),
    "deconv1.fwd": lambda R: run("deconv1.fwd", plan.deconv_k4s1p0_plan("d1", 0, 256, 128, 5), "fwd", R, (5, 5, 256), (8, 8, 128)),
    "conv2.fwd": lambda R: run("conv2.fwd", plan.conv_s2_plan("c2", 0, 32, 64, 32), "fwd", R // 4, (32, 32, 32), (16, 16, 64)),
    "conv3.fwd": lambda R: run("conv3.fwd", plan.conv_s2_plan("c3", 0, 64, 128, 16), "fwd", R // 4, (16, 16, 64), (8, 8, 128)),
    "conv4.fwd": lambda R: run("conv4.fwd", plan.conv_k4s1p0_plan("c4", 0, 128, 256, 8), "fwd", R // 4, (8, 8, 128), (5, 5, 256)),
    "conv2.dgrad": lambda R: run("conv2.dgrad", plan.conv_s2_plan("c2", 0, 32, 64, 32), "dgrad", R // 4, (16, 16, 64), (32, 32, 32)),
    "conv3.dgrad": lambda R: run("conv3.dgrad", plan.conv_s2_plan("c3", 0, 64, 128, 16), "dgrad", R // 4, (8, 8, 128), (16, 16, 64)),
    "deconv3.dgrad": lambda R: run("deconv3.dgrad", plan.deconv_s2_plan("d3", 0, 64, 32, 16), "dgrad", R, (32, 32, 32), (16, 16, 64)),
    "deconv1.wgrad": lambda R: run_wgrad("deconv1.wgrad", plan.deconv_k4s1p0_plan("d1", 0, 256, 128, 5), R, (8, 8, 128), (5, 5, 256)),
    "deconv2.wgrad": lambda R: run_wgrad("deconv2.wgrad", plan.deconv_s2_plan("d2", 0, 128, 64, 8), R, (16, 16, 64), (8, 8, 128)),
    "deconv3.wgrad": lambda R: run_wgrad("deconv3.wgrad", plan.deconv_s2_plan("d3", 0, 64, 32, 16), R, (32, 32, 32), (16, 16, 64)),
    "deconv4.wgrad": lambda R: run_wgrad("deconv4.wgrad", plan.deconv_out_plan("d4", 0), R, (66, 66, 4), (32, 32, 32)),
    "conv2.wgrad": lambda R: run_wgrad("conv2.wgrad", plan.conv_s2_plan("c2", 0, 32, 64, 32), R // 4, (32, 32, 32), (16, 16, 64)),
    "conv3.wgrad": lambda R: run_wgrad("conv3.wgrad", plan.conv_s2_plan("c3", 0, 64, 128, 16), R // 4, (16, 16, 64), (8, 8, 128)),
    "conv4.wgrad": lambda R: run_wgrad("conv4.wgrad", plan.conv_k4s1p0_plan("c4", 0, 128, 256, 8), R // 4, (8, 8, 128), (5, 5, 256)),
    "deconv4.dgrad": lambda R: run("deconv4.dgrad", plan.deconv_out_plan("d4", 0), "dgrad", R, (66, 66, 4), (32, 32, 32)),
    "deconv2.dgrad": lambda R: run("deconv2.dgrad", plan.deconv_s2_plan("d2", 0, 128, 64, 8), "dgrad", R, (16, 16, 64), (8, 8, 128)),
    "deconv3.dgrad.pairs": lambda R: run("deconv3.dgrad.pr", plan.deconv_s2_plan("d3", 0, 64, 32, 16), "dgrad", R, (32, 34, 32), (16, 16, 64),
                                         geom=plan.deconv_s2_plan("d3", 0, 64, 32, 16).extra["pair_dgrad"]),
    "deconv3.wgrad.pairs": lambda R: run_wgrad("deconv3.wgrad.pr", plan.deconv_s2_plan("d3", 0, 64, 32, 16), R, (32, 34, 32), (16, 16, 64),
                                               wg=plan.deconv_s2_plan("d3", 0, 64, 32, 16).extra["pair_wgrad"]),
    "deconv1.dgrad": lambda R: run("deconv1.dgrad", plan.deconv_k4s1p0_plan("d1", 0, 256, 128, 5), "dgrad", R, (8, 8, 128), (5, 5, 256)),
    "conv4.dgrad": lambda R: run("conv4.dgrad", plan.conv_k4s1p0_plan("c4", 0, 128, 256, 8), "dgrad", R // 4, (5, 5, 256), (8, 8, 128)),
}

if __name__ == "__main__":
    R = int(os.environ.get("ROWS", 4096))
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        CASES[nm](R)
