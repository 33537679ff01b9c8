"""Input-pipeline parity on the CPU: the oracle restatement of Pillow's 8-bit
antialiased bilinear resize + ToTensor against golden outputs of the real reference transform
(tests/golden/make_golden_resize.py), and the library's HOST-side coefficient tables against the
oracle — all bit-exact (integer / byte arithmetic)."""
import os

import numpy as np
import pytest

from oracle import resize_oracle as ro
from tests.golden.make_golden_resize import CASES, frames_for

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "resize.npz"))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_resize_matches_reference_transform(case):
    name, n, h, w, oh, ow = case
    out = ro.frames_to_tensor(frames_for(name, n, h, w), oh, ow)
    gold = GOLD[name].astype(np.float32) / np.float32(255.0)
    assert out.dtype == np.float32 and out.shape == gold.shape
    assert np.array_equal(out, gold)


def test_oracle_gather_and_identity():
    f = frames_for("same", 5, 64, 64)
    out = ro.frames_to_tensor(f, 64, 64, index=[4, 0, 4])
    assert np.array_equal(out[0], f[4].transpose(2, 0, 1).astype(np.float32) / np.float32(255))
    assert np.array_equal(out[0], out[2]) and not np.array_equal(out[0], out[1])


@pytest.mark.parametrize("sizes", [(256, 256, 64, 64), (200, 300, 64, 64), (50, 70, 64, 64), (64, 64, 64, 64),
                                   (100, 64, 64, 64), (255, 257, 64, 64), (1000, 999, 64, 64), (256, 256, 32, 48)])
def test_host_coefficient_table_matches_oracle(sizes):
    """mmdyn_resize_table (C++, double precision on the host) == precompute_coeffs + normalize_coeffs_8bpc."""
    from mmdyn_b200 import ops
    H, W, oh, ow = sizes
    t = ops.resize_table(H, W, oh, ow).numpy()
    ksx, bx, kx = ro.precompute_coeffs(W, ow)
    ksy, by, ky = ro.precompute_coeffs(H, oh)
    exp = np.concatenate([[ksx, ksy, H, W, oh, ow, 0, 0], bx.ravel(), kx.ravel(), by.ravel(), ky.ravel()])
    assert np.array_equal(t, exp.astype(np.int32))
    # every coefficient row sums to 2^22 up to the per-tap rounding
    assert np.all(np.abs(kx.sum(1) - (1 << 22)) <= kx.shape[1]) and np.all(np.abs(ky.sum(1) - (1 << 22)) <= ky.shape[1])


def test_resize_table_rejects_bad_sizes():
    from mmdyn_b200 import lib, ops
    with pytest.raises(ValueError):
        ops.resize_table(0, 64, 64, 64)
    with pytest.raises(lib.MmdynError):
        ops.resize_table(64 * 40, 64, 64, 64)  # more than 31x down-scaling
