"""CPU emulation of the *semantics* of mmdyn_igemm / mmdyn_wgrad (test infrastructure only).

Executes a GemmGeom / WgradGeom exactly as include/mmdyn_b200.h specifies it, in fp64 numpy, so
that the host-side planning (tap tables, weight packing) can be validated against torch's own
conv2d / conv_transpose2d without a GPU, and so GPU failures isolate to kernel mechanics."""
import numpy as np


def window_view(buf, n_img, IH, IW, C, pix_stride, row_stride, img_stride):
    """Strided (possibly overlapping) view [n_img, IH, IW, C] of a flat operand buffer, exactly the
    addressing of the 4-D tensor map built in mmdyn_igemm / mmdyn_wgrad (strides in elements)."""
    flat = np.ascontiguousarray(buf).reshape(-1)
    last = (n_img - 1) * img_stride + (IH - 1) * row_stride + (IW - 1) * pix_stride + C
    assert last <= flat.size, "operand view reads past the end of the buffer"
    e = flat.itemsize
    return np.lib.stride_tricks.as_strided(flat, (n_img, IH, IW, C), (img_stride * e, row_stride * e, pix_stride * e, e),
                                           writeable=False)


def pack(flat_params, idx):
    out = np.zeros(idx.shape, np.float64)
    m = idx >= 0
    out[m] = flat_params[idx[m]]
    return out


def igemm(geom, A, Wp, n_img, bias=None):
    """A: [n_img, IH, IW, a_pix_stride]; Wp: [n_phases*N, K]. Returns NHWC [n_img, OH, OW, ldc]
    (out_mode 0/1) or NCHW [n_img, 3, OH, OW] (out_mode 3)."""
    g = geom
    OYv = g.P // g.OXv
    if g.a_row_stride or g.a_img_stride or g.a_pix_stride < g.Cin:
        A = window_view(A, n_img, g.IH, g.IW, g.Cin, g.a_pix_stride, g.a_row_stride or g.a_pix_stride * g.IW,
                        g.a_img_stride or (g.a_row_stride or g.a_pix_stride * g.IW) * g.IH)
    if g.out_mode == 3:
        out = np.zeros((n_img, 3, g.OH, g.OW))
    else:
        out = np.zeros((n_img, g.OH, g.OW, g.ldc))
    for ph in range(g.n_phases):
        W = Wp[ph * g.N:(ph + 1) * g.N]
        for yv in range(OYv):
            for xv in range(g.OXv):
                acc = np.zeros((n_img, g.N))
                for t in range(g.ntaps):
                    iy = yv * g.s_in + g.tap_dy[ph][t]
                    ix = xv * (g.s_in_x or g.s_in) + g.tap_dx[ph][t]
                    if 0 <= iy < g.IH and 0 <= ix < g.IW:
                        acc += A[:, iy, ix, :g.Cin] @ W[:, t * g.Cin:(t + 1) * g.Cin].T
                if bias is not None:
                    acc += bias[None, :]
                if g.out_mode == 4:
                    C = g.ldc
                    for p2 in range(2):
                        for q2 in range(2):
                            out[:, 2 * yv + p2, 2 * xv + q2, :C] = acc[:, (p2 * 2 + q2) * C:(p2 * 2 + q2 + 1) * C]
                elif g.out_mode == 3:
                    for p2 in range(2):
                        for q2 in range(2):
                            for c in range(3):
                                out[:, c, 2 * yv + p2, 2 * xv + q2] = acc[:, (p2 * 2 + q2) * 3 + c]
                else:
                    oy = yv * g.s_out + g.off_y[ph]
                    ox = xv * g.s_out + g.off_x[ph]
                    out[:, oy, ox, :g.N] = acc
    return out


def wgrad(geom, G, Nat, n_img):
    """G: [n_img, IH, IW, g_pix_stride]; Nat: [n_img, P, nat_stride]. Returns dW [Cn, K]."""
    g = geom
    OYv = g.P // g.OXv
    if g.g_row_stride or g.g_img_stride or g.g_pix_stride < g.Cg:
        G = window_view(G, n_img, g.IH, g.IW, g.Cg, g.g_pix_stride, g.g_row_stride or g.g_pix_stride * g.IW,
                        g.g_img_stride or (g.g_row_stride or g.g_pix_stride * g.IW) * g.IH)
    dW = np.zeros((g.Cn, g.K))
    for yv in range(OYv):
        for xv in range(g.OXv):
            nat = Nat[:, yv * g.OXv + xv, :g.Cn]
            for t in range(g.ntaps):
                iy = yv * g.s_in + g.tap_dy[t]
                ix = xv * (g.s_in_x or g.s_in) + g.tap_dx[t]
                if 0 <= iy < g.IH and 0 <= ix < g.IW:
                    dW[:, t * g.Cg:(t + 1) * g.Cg] += nat.T @ G[:, iy, ix, :g.Cg]
    return dW


def unpack_add(dWp, idx, n_params):
    flat = np.zeros(n_params)
    m = idx >= 0
    np.add.at(flat, idx[m], dWp[m])
    return flat
