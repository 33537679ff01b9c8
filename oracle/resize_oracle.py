"""TEST INFRASTRUCTURE ONLY — CPU restatement of the image transform in front of the cnn-vae / cnn-mvae
step: `transforms.Compose([Resize(input_size), ToTensor()])`
(mmdyn/pytorch/utils/datasets.py:23-31), applied per frame in `VisuoTactileDataset._parse_list_data`
(datasets.py:382-392) to uint8 HxWx3 arrays via `Image.fromarray`.

The arithmetic lives in a third-party dependency that is not vendored in the reference tree:
Pillow (un-pinned in the reference's setup.py; 12.2.0 in this image), `src/libImaging/Resample.c`:
torchvision's `Resize` on a PIL image calls `Image.resize(size, BILINEAR)` = a separable, antialiased
triangle filter evaluated in 8-bit fixed point (`precompute_coeffs`, `normalize_coeffs_8bpc`,
`ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`; PRECISION_BITS = 32 - 8 - 2 = 22),
horizontal pass first, each pass rounding to uint8; a pass whose input and output sizes agree is
skipped.  `ToTensor` then yields CHW float32 = uint8 / 255.

Parity is pinned: tests/test_resize_cpu.py checks this restatement bit for bit against golden outputs
produced by the real PIL + torchvision pipeline in the build container (tests/golden/make_golden_resize.py).
Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may import this module."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bilinear(x):
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs (box = the whole axis, filter = BILINEAR, support 1.0) followed by
    normalize_coeffs_8bpc.  Returns (ksize, bounds [out,2] int32, kk [out,ksize] int32)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)          # C (int) cast: truncation toward zero
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size) - xmin
        w = [_bilinear((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass_axis1(img, out_size):
    """8bpc resample along axis 1 of img [R, in, C] uint8 -> [R, out, C] uint8."""
    _, bounds, kk = precompute_coeffs(img.shape[1], out_size)
    out = np.empty((img.shape[0], out_size, img.shape[2]), np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, xmax = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, xmin:xmin + xmax, :], kk[xx, :xmax].astype(np.int64),
                                                         axes=([1], [0]))
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_u8(img, out_h, out_w):
    """PIL Image.fromarray(img).resize((out_w, out_h), BILINEAR) for img [H, W, C] uint8."""
    assert img.dtype == np.uint8 and img.ndim == 3
    if img.shape[1] != out_w:
        img = _pass_axis1(img, out_w)                                   # horizontal pass first
    if img.shape[0] != out_h:
        img = _pass_axis1(img.transpose(1, 0, 2), out_h).transpose(1, 0, 2)
    return np.ascontiguousarray(img)


def frames_to_tensor(frames, out_h, out_w, index=None):
    """Resize + ToTensor for a stack of frames [N, H, W, 3] uint8 (optionally gathered by `index`):
    float32 [n, 3, out_h, out_w] = resized uint8 / 255 (torchvision ToTensor: .float().div(255))."""
    idx = range(frames.shape[0]) if index is None else [int(i) for i in index]
    out = np.empty((len(idx), frames.shape[3], out_h, out_w), np.float32)
    for o, i in enumerate(idx):
        r = resize_u8(frames[i], out_h, out_w)
        out[o] = r.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
    return out
