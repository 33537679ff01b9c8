"""Layer geometry -> implicit-GEMM descriptors and weight pack/unpack index maps.

Pure host logic (numpy, no CUDA): for every Conv2d / ConvTranspose2d / Linear of the reference
model (mmdyn/pytorch/models/vae.py:197-216, 263-279) this module derives

  * the tap tables / virtual grids consumed by `mmdyn_igemm` for the forward pass and for the
    input gradient (dgrad), and by `mmdyn_wgrad` for the weight gradient;
  * int32 index maps that gather the fp32 torch-layout parameters (living in one flat arena)
    into the fp16 K-contiguous operand matrices the kernels read, and scatter packed weight
    gradients back.

A stride-2 ConvTranspose2d (k4, p1) is decomposed into its 4 sub-pixel phases (each a 2x2-tap
stride-1 convolution); the dgrad of a stride-2 Conv2d is the same thing with the channel roles
swapped.  The k4/s1/p0 layers (8->5 conv, 5->8 deconv) use "pixel-major" tiles so that taps that
fall outside the 5x5 map are skipped instead of multiplied by zeros.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import os

import numpy as np

MAX_TAPS = 16
# run the 4 sub-pixel phases of stride-2 (de)convs as one GEMM (see _merged_geom); MMDYN_MERGE_PHASES=0
# keeps them as 4 phase GEMMs of 4 taps each (no zero-weight MMAs, 16 operand boxes instead of 9)
MERGE_PHASES = os.environ.get("MMDYN_MERGE_PHASES", "1") != "0"
# merged 3x3-tap layers on the patch-reuse kernel (one activation box per filter column, live weight blocks
# only); MMDYN_NO_PATCH=1 keeps them on the generic one-box-per-tap kernel (A/B measurements)
USE_PATCH = os.environ.get("MMDYN_NO_PATCH") is None
LOGIT_CP = 4  # channels per pixel of the stored logit gradients (3 used): fp16 NHWC4 with a one-pixel zero border
PATCH_SPLIT = os.environ.get("MMDYN_NO_PATCH_SPLIT") is None  # deconv2.fwd / conv3.dgrad on the phase-split patch kernel


@dataclass
class GemmGeom:
    P: int
    OXv: int
    IH: int
    IW: int
    Cin: int
    s_in: int
    tap_dy: List[List[int]]   # [phase][tap]
    tap_dx: List[List[int]]
    N: int
    OH: int
    OW: int
    s_out: int
    off_y: List[int]
    off_x: List[int]
    ldc: int
    row_mode: int = 0
    out_mode: int = 0
    a_pix_stride: Optional[int] = None
    a_row_stride: int = 0     # 0 = dense; see include/mmdyn_b200.h (overlapping-window operands)
    a_img_stride: int = 0
    s_in_x: int = 0           # input stride along x when it differs from s_in (0 = s_in)
    patch: int = 0            # 1: merged 3x3-tap layer -> shared-memory patch reuse kernel (mmdyn_igemm patch_mode)

    @property
    def ntaps(self):
        return len(self.tap_dy[0])

    @property
    def n_phases(self):
        return len(self.tap_dy)

    @property
    def K(self):
        return self.ntaps * self.Cin

    @property
    def block_n(self):
        return min(self.N, 256)

    def __post_init__(self):
        if self.a_pix_stride is None:
            self.a_pix_stride = self.Cin
        assert self.K % 64 == 0, (self.ntaps, self.Cin)
        assert self.N % self.block_n == 0 and self.block_n in (16, 32, 64, 128, 256), self.N
        assert self.ntaps <= MAX_TAPS


@dataclass
class WgradGeom:
    P: int
    OXv: int
    IH: int
    IW: int
    Cg: int
    s_in: int
    tap_dy: List[int]
    tap_dx: List[int]
    Cn: int
    g_pix_stride: Optional[int] = None
    nat_stride: Optional[int] = None
    g_row_stride: int = 0
    g_img_stride: int = 0
    s_in_x: int = 0

    @property
    def ntaps(self):
        return len(self.tap_dy)

    @property
    def K(self):
        return self.ntaps * self.Cg

    def __post_init__(self):
        if self.g_pix_stride is None:
            self.g_pix_stride = self.Cg
        if self.nat_stride is None:
            self.nat_stride = self.Cn
        assert self.K % 128 == 0 or (self.Cg == 16 and self.ntaps == 4), (self.ntaps, self.Cg)


@dataclass
class LayerPlan:
    """Everything the engine needs for one weight-bearing layer."""
    name: str
    kind: str
    fwd: Optional[GemmGeom]
    idx_fwd: Optional[np.ndarray]        # [n_phases*N, K] -> arena index (-1 = zero)
    dgrad: Optional[GemmGeom]
    idx_dgrad: Optional[np.ndarray]
    wgrad: Optional[WgradGeom]
    idx_wgrad: Optional[np.ndarray]      # [Cn, K_w] -> arena index (-1 = padding)
    bias_idx: Optional[np.ndarray] = None  # [N] arena indices of the (permuted) bias
    extra: dict = field(default_factory=dict)


# sub-pixel phase tables for k4 / s2 / p1: phase parity -> [(kernel index, input offset)]
#   out = 2*in - 1 + k  =>  for out = 2*v + ph: k = 1 -> in = v (ph 0), k = 3 -> in = v-1 (ph 0),
#                                                k = 0 -> in = v+1 (ph 1), k = 2 -> in = v (ph 1)
PH_TAPS = {0: [(1, 0), (3, -1)], 1: [(0, 1), (2, 0)]}


def _conv_taps(k, pad):
    dy = [kh - pad for kh in range(k) for _ in range(k)]
    dx = [kw - pad for _ in range(k) for kw in range(k)]
    return dy, dx


def _phase_taps():
    tdy, tdx, tk = [], [], []
    for ph in (0, 1):
        for pw in (0, 1):
            dy, dx, ks = [], [], []
            for (kh, oy) in PH_TAPS[ph]:
                for (kw, ox) in PH_TAPS[pw]:
                    dy.append(oy)
                    dx.append(ox)
                    ks.append((kh, kw))
            tdy.append(dy)
            tdx.append(dx)
            tk.append(ks)
    return tdy, tdx, tk


# merged sub-pixel phases: union of the 4 phases' taps (3x3 input offsets); phase parity ->
# {input offset: kernel index}
KOF = {0: {0: 1, -1: 3}, 1: {1: 0, 0: 2}}
TAPS3 = [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)]


def _merged_geom(Hv, Cg, Cn_out, widx_fn):
    """One GEMM for all 4 sub-pixel phases of a k4/s2/p1 transposed conv (or of the dgrad of a k4/s2/p1
    conv): rows = the Hv x Hv low-resolution grid, K = 9 taps x Cg gathered channels, N = 4 x Cn_out
    (n = (ph*2+pw)*Cn_out + c), out_mode 4 scatters the 2x2 output pixels.  9 operand boxes per 128
    rows instead of 16 — these layers are operand-feed bound, the extra (zero-weight) MACs are free."""
    geom = GemmGeom(P=Hv * Hv, OXv=Hv, IH=Hv, IW=Hv, Cin=Cg, s_in=1, tap_dy=[[t[0] for t in TAPS3]],
                    tap_dx=[[t[1] for t in TAPS3]], N=4 * Cn_out, OH=2 * Hv, OW=2 * Hv, s_out=2, off_y=[0],
                    off_x=[0], ldc=Cn_out, out_mode=4,
                    patch=int(USE_PATCH and Hv in (8, 16, 32) and (Cg % 64 == 0) and
                              ((4 * Cn_out in (64, 128) and 16 * Cn_out * 128 * (Cg // 64) <= 64 * 1024)  # live blocks resident
                               or (PATCH_SPLIT and 4 * Cn_out == 256 and Cg == 128))))  # phase-split: half of them per CTA
    idx = np.full((4 * Cn_out, 9 * Cg), -1, np.int32)
    for ph in (0, 1):
        for pw in (0, 1):
            for t_, (a, b) in enumerate(TAPS3):
                if a in KOF[ph] and b in KOF[pw]:
                    n_, g_ = np.meshgrid(np.arange(Cn_out), np.arange(Cg), indexing="ij")
                    r0 = (ph * 2 + pw) * Cn_out
                    idx[r0:r0 + Cn_out, t_ * Cg:(t_ + 1) * Cg] = widx_fn(n_, g_, KOF[ph][a], KOF[pw][b])
    return geom, idx


def conv_s2_plan(name, w_off, Cin, Cout, H):
    """nn.Conv2d(Cin, Cout, 4, 2, 1, bias=False) on HxH (vae.py:200,203). weight [Cout][Cin][4][4]."""
    Ho = H // 2

    def widx(co, ci, kh, kw):
        return w_off + ((co * Cin + ci) * 4 + kh) * 4 + kw

    dy, dx = _conv_taps(4, 1)
    fwd = GemmGeom(P=Ho * Ho, OXv=Ho, IH=H, IW=H, Cin=Cin, s_in=2, tap_dy=[dy], tap_dx=[dx], N=Cout,
                   OH=Ho, OW=Ho, s_out=1, off_y=[0], off_x=[0], ldc=Cout)
    co, t, ci = np.meshgrid(np.arange(Cout), np.arange(16), np.arange(Cin), indexing="ij")
    idx_fwd = widx(co, ci, t // 4, t % 4).reshape(Cout, 16 * Cin).astype(np.int32)

    # dgrad: 4 phases on the (Ho x Ho) grid of dY, output = dX (H x H x Cin)
    tdy, tdx, tk = _phase_taps()
    dg = GemmGeom(P=Ho * Ho, OXv=Ho, IH=Ho, IW=Ho, Cin=Cout, s_in=1, tap_dy=tdy, tap_dx=tdx, N=Cin,
                  OH=H, OW=H, s_out=2, off_y=[0, 0, 1, 1], off_x=[0, 1, 0, 1], ldc=Cin)
    idx_dg = np.full((4 * Cin, 4 * Cout), -1, np.int32)
    for p in range(4):
        for t_, (kh, kw) in enumerate(tk[p]):
            ci_, co_ = np.meshgrid(np.arange(Cin), np.arange(Cout), indexing="ij")
            idx_dg[p * Cin:(p + 1) * Cin, t_ * Cout:(t_ + 1) * Cout] = widx(co_, ci_, kh, kw)
    if MERGE_PHASES and (9 * Cout) % 64 == 0 and 4 * Cin <= 256 and Cin % 16 == 0:
        dg, idx_dg = _merged_geom(Ho, Cout, Cin, lambda ci_, co_, kh, kw: widx(co_, ci_, kh, kw))
    wg = WgradGeom(P=Ho * Ho, OXv=Ho, IH=H, IW=H, Cg=Cin, s_in=2, tap_dy=dy, tap_dx=dx, Cn=Cout)
    return LayerPlan(name, "conv_s2", fwd, idx_fwd, dg, idx_dg, wg, idx_fwd, extra={"macs": Ho * Ho * Cout * Cin * 16})


def conv_k4s1p0_plan(name, w_off, Cin, Cout, H):
    """nn.Conv2d(128, 256, 4, 1, 0, bias=False) 8x8 -> 5x5 (vae.py:206)."""
    Ho = H - 3

    def widx(co, ci, kh, kw):
        return w_off + ((co * Cin + ci) * 4 + kh) * 4 + kw

    dy, dx = _conv_taps(4, 0)
    fwd = GemmGeom(P=Ho * Ho, OXv=Ho, IH=H, IW=H, Cin=Cin, s_in=1, tap_dy=[dy], tap_dx=[dx], N=Cout,
                   OH=Ho, OW=Ho, s_out=1, off_y=[0], off_x=[0], ldc=Cout)
    co, t, ci = np.meshgrid(np.arange(Cout), np.arange(16), np.arange(Cin), indexing="ij")
    idx_fwd = widx(co, ci, t // 4, t % 4).reshape(Cout, 16 * Cin).astype(np.int32)
    # dgrad: dX[iy] = sum_kh dY[iy - kh] w[kh]; grid = HxH, pixel-major tiles skip out-of-range taps
    ndy, ndx = [-v for v in dy], [-v for v in dx]
    dg = GemmGeom(P=H * H, OXv=H, IH=Ho, IW=Ho, Cin=Cout, s_in=1, tap_dy=[ndy], tap_dx=[ndx], N=Cin,
                  OH=H, OW=H, s_out=1, off_y=[0], off_x=[0], ldc=Cin, row_mode=1)
    ci_, t_, co_ = np.meshgrid(np.arange(Cin), np.arange(16), np.arange(Cout), indexing="ij")
    idx_dg = widx(co_, ci_, t_ // 4, t_ % 4).reshape(Cin, 16 * Cout).astype(np.int32)
    wg = WgradGeom(P=Ho * Ho, OXv=Ho, IH=H, IW=H, Cg=Cin, s_in=1, tap_dy=dy, tap_dx=dx, Cn=Cout)
    return LayerPlan(name, "conv_s1", fwd, idx_fwd, dg, idx_dg, wg, idx_fwd, extra={"macs": Ho * Ho * Cout * Cin * 16})


def deconv_k4s1p0_plan(name, w_off, Cin, Cout, H):
    """nn.ConvTranspose2d(256, 128, 4, 1, 0, bias=False) 5x5 -> 8x8 (vae.py:268). weight [Cin][Cout][4][4]."""
    Ho = H + 3

    def widx(ci, co, kh, kw):
        return w_off + ((ci * Cout + co) * 4 + kh) * 4 + kw

    dy, dx = _conv_taps(4, 0)
    ndy, ndx = [-v for v in dy], [-v for v in dx]
    fwd = GemmGeom(P=Ho * Ho, OXv=Ho, IH=H, IW=H, Cin=Cin, s_in=1, tap_dy=[ndy], tap_dx=[ndx], N=Cout,
                   OH=Ho, OW=Ho, s_out=1, off_y=[0], off_x=[0], ldc=Cout, row_mode=1)
    co, t, ci = np.meshgrid(np.arange(Cout), np.arange(16), np.arange(Cin), indexing="ij")
    idx_fwd = widx(ci, co, t // 4, t % 4).reshape(Cout, 16 * Cin).astype(np.int32)
    # dgrad: dIn[iy] = sum_kh dOut[iy + kh] w[kh]  (a plain valid convolution over dOut)
    dg = GemmGeom(P=H * H, OXv=H, IH=Ho, IW=Ho, Cin=Cout, s_in=1, tap_dy=[dy], tap_dx=[dx], N=Cin,
                  OH=H, OW=H, s_out=1, off_y=[0], off_x=[0], ldc=Cin)
    ci_, t_, co_ = np.meshgrid(np.arange(Cin), np.arange(16), np.arange(Cout), indexing="ij")
    idx_dg = widx(ci_, co_, t_ // 4, t_ % 4).reshape(Cin, 16 * Cout).astype(np.int32)
    # wgrad: natural operand = layer input (rows = 5x5 input pixels), gathered = dOut
    wg = WgradGeom(P=H * H, OXv=H, IH=Ho, IW=Ho, Cg=Cout, s_in=1, tap_dy=dy, tap_dx=dx, Cn=Cin)
    return LayerPlan(name, "deconv_s1", fwd, idx_fwd, dg, idx_dg, wg, idx_dg, extra={"macs": H * H * Cout * Cin * 16})


def deconv_s2_plan(name, w_off, Cin, Cout, H):
    """nn.ConvTranspose2d(Cin, Cout, 4, 2, 1, bias=False) HxH -> 2Hx2H (vae.py:271,274)."""
    Ho = 2 * H

    def widx(ci, co, kh, kw):
        return w_off + ((ci * Cout + co) * 4 + kh) * 4 + kw

    tdy, tdx, tk = _phase_taps()
    fwd = GemmGeom(P=H * H, OXv=H, IH=H, IW=H, Cin=Cin, s_in=1, tap_dy=tdy, tap_dx=tdx, N=Cout,
                   OH=Ho, OW=Ho, s_out=2, off_y=[0, 0, 1, 1], off_x=[0, 1, 0, 1], ldc=Cout)
    idx_fwd = np.full((4 * Cout, 4 * Cin), -1, np.int32)
    for p in range(4):
        for t_, (kh, kw) in enumerate(tk[p]):
            co_, ci_ = np.meshgrid(np.arange(Cout), np.arange(Cin), indexing="ij")
            idx_fwd[p * Cout:(p + 1) * Cout, t_ * Cin:(t_ + 1) * Cin] = widx(ci_, co_, kh, kw)
    if MERGE_PHASES and (9 * Cin) % 64 == 0 and 4 * Cout <= 256 and Cout % 16 == 0:
        fwd, idx_fwd = _merged_geom(H, Cin, Cout, lambda co_, ci_, kh, kw: widx(ci_, co_, kh, kw))
    dy, dx = _conv_taps(4, 1)
    dg = GemmGeom(P=H * H, OXv=H, IH=Ho, IW=Ho, Cin=Cout, s_in=2, tap_dy=[dy], tap_dx=[dx], N=Cin,
                  OH=H, OW=H, s_out=1, off_y=[0], off_x=[0], ldc=Cin)
    ci_, t_, co_ = np.meshgrid(np.arange(Cin), np.arange(16), np.arange(Cout), indexing="ij")
    idx_dg = widx(ci_, co_, t_ // 4, t_ % 4).reshape(Cin, 16 * Cout).astype(np.int32)
    wg = WgradGeom(P=H * H, OXv=H, IH=Ho, IW=Ho, Cg=Cout, s_in=2, tap_dy=dy, tap_dx=dx, Cn=Cin)
    extra = {"macs": H * H * Cout * Cin * 16}
    if Cout == 32 and H >= 16 and (H & (H - 1)) == 0:
        # 32-channel gradient (64-byte pixels): the same contraction over 128-byte pixel PAIRS.  With the gradient stored
        # with one extra pixel on each side of every image row ([img][Ho][Ho + 2][32], mmdyn_bn_bwd_apply_padded) the four
        # x-taps kw = 0..3 of output pixel x are padded pixels 2x .. 2x + 3 = pairs x and x + 1: 8 taps (kh, pair) of 64
        # "channels" instead of 16 taps of 32 — half the TMA row requests for the same bytes, and the K order (kh, kw, c)
        # of the packed weights is unchanged.  Rows above / below the image come from the TMA zero fill as before.
        pdy = [kh - 1 for kh in range(4) for _ in range(2)]
        pdx = [p for _ in range(4) for p in range(2)]
        row, img = (Ho + 2) * Cout, Ho * (Ho + 2) * Cout
        extra["pair_dgrad"] = GemmGeom(P=H * H, OXv=H, IH=Ho, IW=H + 1, Cin=2 * Cout, s_in=2, s_in_x=1, tap_dy=[pdy],
                                       tap_dx=[pdx], N=Cin, OH=H, OW=H, s_out=1, off_y=[0], off_x=[0], ldc=Cin,
                                       a_pix_stride=2 * Cout, a_row_stride=row, a_img_stride=img)
        extra["pair_wgrad"] = WgradGeom(P=H * H, OXv=H, IH=Ho, IW=H + 1, Cg=2 * Cout, s_in=2, s_in_x=1, tap_dy=pdy,
                                        tap_dx=pdx, Cn=Cin, g_pix_stride=2 * Cout, g_row_stride=row, g_img_stride=img)
    return LayerPlan(name, "deconv_s2", fwd, idx_fwd, dg, idx_dg, wg, idx_dg, extra=extra)


def deconv_out_plan(name, w_off, Cin=32, Cout=3, H=32):
    """nn.ConvTranspose2d(32, 3, 4, 2, 1, bias=False) 32x32 -> 64x64 logits (vae.py:277).

    Forward: the 4 sub-pixel phases are merged into the N dimension (n = (ph*2+pw)*3 + co, 12 of
    16 used) over the union of their taps (3x3 + one zero-weight dummy tap so K = 10*32 = 320);
    the epilogue writes fp32 NCHW planes.  Backward works on dlogits stored NHWC with LOGIT_CP = 4
    channels per pixel (3 used)."""
    Ho = 2 * H
    CP = LOGIT_CP  # padded gradient channels

    def widx(ci, co, kh, kw):
        return w_off + ((ci * Cout + co) * 4 + kh) * 4 + kw

    taps = [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)] + [(0, 0)]
    kof = {0: {0: 1, -1: 3}, 1: {1: 0, 0: 2}}  # phase parity -> {input offset: kernel index}
    fwd = GemmGeom(P=H * H, OXv=H, IH=H, IW=H, Cin=Cin, s_in=1, tap_dy=[[t[0] for t in taps]],
                   tap_dx=[[t[1] for t in taps]], N=16, OH=Ho, OW=Ho, s_out=2, off_y=[0], off_x=[0],
                   ldc=0, out_mode=3, patch=int(USE_PATCH and H == 32 and Cin == 32))
    idx_fwd = np.full((16, len(taps) * Cin), -1, np.int32)
    for ph in (0, 1):
        for pw in (0, 1):
            for co in range(Cout):
                n = (ph * 2 + pw) * 3 + co
                for t_, (a, b) in enumerate(taps[:9]):
                    if a in kof[ph] and b in kof[pw]:
                        idx_fwd[n, t_ * Cin:(t_ + 1) * Cin] = widx(np.arange(Cin), co, kof[ph][a], kof[pw][b])
    # Backward operand = dlogits, NHWC with CP = 4 channels per pixel (8 bytes: 1.33x the 3 algorithmic
    # channels), stored with a one-pixel zero border: [img][Ho+2][Ho+2][4].  The 4 x-taps of an output
    # pixel (padded columns 2x .. 2x+3) are 32 contiguous bytes, so they are read as ONE tap of a "window
    # pixel" with 4*CP = 16 channels.  TMA strides are multiples of 16 bytes, so the x dimension counts
    # pixel PAIRS (a_pix_stride = 2*CP = 8 elements < Cin = 16): window x covers pairs x, x+1 and advances one
    # pair per output pixel (s_in_x = 1) while rows advance two per output row (s_in = 2).  4 operand boxes
    # of 32-byte rows (K = 64, one k-block) per tile, and no out-of-image taps.  K order (kh, kw, c).
    Hp = Ho + 2
    wdy, wdx = [0, 1, 2, 3], [0, 0, 0, 0]
    dg = GemmGeom(P=H * H, OXv=H, IH=Hp, IW=H, Cin=4 * CP, s_in=2, s_in_x=1, tap_dy=[wdy], tap_dx=[wdx], N=Cin,
                  OH=H, OW=H, s_out=1, off_y=[0], off_x=[0], ldc=Cin, a_pix_stride=2 * CP, a_row_stride=Hp * CP,
                  a_img_stride=Hp * Hp * CP)
    idx_dg = np.full((Cin, 16 * CP), -1, np.int32)
    for t_ in range(16):
        for co in range(Cout):
            idx_dg[:, t_ * CP + co] = widx(np.arange(Cin), co, t_ // 4, t_ % 4)
    wg = WgradGeom(P=H * H, OXv=H, IH=Hp, IW=H, Cg=4 * CP, s_in=2, s_in_x=1, tap_dy=wdy, tap_dx=wdx, Cn=Cin,
                   g_pix_stride=2 * CP, g_row_stride=Hp * CP, g_img_stride=Hp * Hp * CP)
    return LayerPlan(name, "deconv_out", fwd, idx_fwd, dg, idx_dg, wg, idx_dg, extra={"macs": H * H * Cout * Cin * 16})


def linear_plan(name, w_offs, b_offs, K, Ns, k_perm=None, n_perm=None, ld=None):
    """One or several nn.Linear(K, N_i) sharing their input, concatenated along N
    (vae.py:211, 215-216, 264).  k_perm[k'] / n_perm[n'] give the torch index of packed index."""
    N = int(sum(Ns))
    ld = K if ld is None else int(ld)  # row pitch of the torch weights: > K when extra input columns
    #                                    (the CVAE condition, vae.py:231-237, 286-291) are handled elsewhere
    kp = np.arange(K) if k_perm is None else np.asarray(k_perm)
    rows, bias = [], []
    for w_off, b_off, n_i in zip(w_offs, b_offs, Ns):
        npm = np.arange(n_i) if n_perm is None else np.asarray(n_perm)
        rows.append(w_off + npm[:, None] * ld + kp[None, :])
        bias.append(b_off + npm)
    idx_fwd = np.concatenate(rows, 0).astype(np.int32)        # [N][K]
    bias_idx = np.concatenate(bias, 0).astype(np.int32)
    fwd = GemmGeom(P=1, OXv=1, IH=1, IW=1, Cin=K, s_in=1, tap_dy=[[0]], tap_dx=[[0]], N=N, OH=1, OW=1,
                   s_out=1, off_y=[0], off_x=[0], ldc=N)
    dg = GemmGeom(P=1, OXv=1, IH=1, IW=1, Cin=N, s_in=1, tap_dy=[[0]], tap_dx=[[0]], N=K, OH=1, OW=1,
                  s_out=1, off_y=[0], off_x=[0], ldc=K)
    idx_dg = np.ascontiguousarray(idx_fwd.T)                   # [K][N]
    wg = WgradGeom(P=1, OXv=1, IH=1, IW=1, Cg=K, s_in=1, tap_dy=[0], tap_dx=[0], Cn=N)
    return LayerPlan(name, "linear", fwd, idx_fwd, dg, idx_dg, wg, idx_fwd, bias_idx=bias_idx, extra={"macs": K * N})


LO_BIT = 1 << 30  # pack index flag: low part of the two-term fp16 split of that weight (mmdyn_pack_f16)


def linear_split_plan(name, w_offs, b_offs, K, Ns):
    """nn.Linear(K, N_i) layers of the fp32 pose expert (vae.py:14-19, 118-123) on the fp16 tensor cores at fp32
    accuracy: every operand is split x = hi + lo and the GEMMs run over the contraction dimension tripled,
    activations as [hi | lo | hi], weights / gradients as [hi | hi | lo], so that one fp32-accumulating GEMM sums
    hi*hi + lo*hi + hi*lo (mmdyn_split_f16).  Forward: rows [M][3K] x packed W [N][3K]; dgrad: rows [M][3N] x
    packed W^T [K][3N] (the gradient rows are mode 1, so W^T is packed [hi | lo | hi]); wgrad: the split rows seen as
    [3M][K] and [3M][N] matrices pair up segment by segment, dW lands in the torch layout [N][K] directly."""
    N = int(sum(Ns))
    rows, bias = [], []
    for w_off, b_off, n_i in zip(w_offs, b_offs, Ns):
        rows.append(w_off + np.arange(n_i)[:, None] * K + np.arange(K)[None, :])
        bias.append(b_off + np.arange(n_i))
    w = np.concatenate(rows, 0).astype(np.int64)                    # [N][K] arena indices
    idx_fwd = np.concatenate([w, w, w | LO_BIT], 1).astype(np.int32)          # [N][3K]   hi | hi | lo
    wt = np.ascontiguousarray(w.T)                                   # [K][N]
    idx_dg = np.concatenate([wt, wt | LO_BIT, wt], 1).astype(np.int32)        # [K][3N]   hi | lo | hi
    fwd = GemmGeom(P=1, OXv=1, IH=1, IW=1, Cin=3 * K, s_in=1, tap_dy=[[0]], tap_dx=[[0]], N=N, OH=1, OW=1,
                   s_out=1, off_y=[0], off_x=[0], ldc=N)
    dg = GemmGeom(P=1, OXv=1, IH=1, IW=1, Cin=3 * N, s_in=1, tap_dy=[[0]], tap_dx=[[0]], N=K, OH=1, OW=1,
                  s_out=1, off_y=[0], off_x=[0], ldc=K)
    wgs = [WgradGeom(P=1, OXv=1, IH=1, IW=1, Cg=K, s_in=1, tap_dy=[0], tap_dx=[0], Cn=int(n_i), nat_stride=N) for n_i in Ns]
    return LayerPlan(name, "linear_split", fwd, idx_fwd, dg, idx_dg, None, None,
                     bias_idx=np.concatenate(bias, 0).astype(np.int32),
                     extra={"macs": 3 * K * N, "wgrads": wgs, "w_offs": list(w_offs), "Ns": [int(n) for n in Ns], "K": K})


def conv1_plan(name, w_off):
    """nn.Conv2d(3, 32, 4, 2, 1, bias=False) on the fp32 NCHW input (vae.py:198): packed [32][64],
    k = ci*16 + kh*4 + kw (the torch layout), 48 used."""
    idx = np.full((32, 64), -1, np.int32)
    idx[:, :48] = w_off + np.arange(32)[:, None] * 48 + np.arange(48)[None, :]
    return LayerPlan(name, "conv1", None, idx, None, None, None, None, extra={"macs": 1024 * 32 * 48})


def conv1_wgrad_plan(w_off):
    """Weight gradient of the first conv through the generic tensor-core wgrad kernel: the fp32 NCHW
    input is repacked once per backward into NHWC fp16 with LOGIT_CP = 4 channels per pixel (3 used) and a
    one-pixel zero border, so dW[co][(kh*4 + kw)*4 + c] = sum_pix dRaw[pix][co] * xp[2*oy + kh][2*ox + kw][c]:
    the same 4-pixel-window operand as the logits layer's weight gradient (deconv_out_plan), K = 64."""
    CP, Hp = LOGIT_CP, 66
    wg = WgradGeom(P=32 * 32, OXv=32, IH=Hp, IW=32, Cg=4 * CP, s_in=2, s_in_x=1, tap_dy=[0, 1, 2, 3],
                   tap_dx=[0, 0, 0, 0], Cn=32, g_pix_stride=2 * CP, g_row_stride=Hp * CP, g_img_stride=Hp * Hp * CP)
    idx = np.full((32, 16 * CP), -1, np.int32)
    for t_ in range(16):
        for c in range(3):
            idx[:, t_ * CP + c] = w_off + (np.arange(32) * 3 + c) * 16 + t_
    return wg, idx


def nhwc_perm(C, H, W):
    """perm[k'] = torch flat index (c*H*W + h*W + w) of NHWC flat index k' = (h*W + w)*C + c."""
    hw, c = np.meshgrid(np.arange(H * W), np.arange(C), indexing="ij")
    return (c * (H * W) + hw).reshape(-1)


def tile_images(geom: GemmGeom) -> int:
    """Images per 128-row tile of mmdyn_igemm's TMA path (same rule as the launcher, csrc/igemm.cu): a box of
    OXv x bh pixels x bn images, or one pixel x 128 images.  Fused BatchNorm statistics need whole tiles per group."""
    OYv, bw = geom.P // geom.OXv, geom.OXv
    if geom.row_mode != 0 or geom.P <= 1 or bw > 128 or 128 % bw:
        return 128
    bh = min(128 // bw, OYv)
    if OYv % bh or bh & (bh - 1) or 128 % (bw * bh):
        return 128
    return 128 // (bw * bh)


def choose_ksplit(geom: GemmGeom, n_img: int, sm_count: int = 148):
    """Split K across CTAs when the output grid alone cannot fill the machine."""
    if geom.row_mode == 1:
        return 1
    rows = n_img * geom.P
    ctas = ((rows + 127) // 128) * (geom.N // geom.block_n) * geom.n_phases
    kb = geom.K // 64
    if ctas >= sm_count or kb < 8:
        return 1
    return int(max(1, min(kb // 4, (2 * sm_count + ctas - 1) // ctas)))


def choose_row_splits(geom: WgradGeom, n_img: int, sm_count: int = 148):
    cn_tile = min(geom.Cn, 256)
    base = ((geom.K + 127) // 128) * (geom.Cn // cn_tile)
    steps = (n_img * geom.P + 63) // 64
    want = (2 * sm_count + base - 1) // base
    return int(max(1, min(steps, want)))
