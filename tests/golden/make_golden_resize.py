"""Generates tests/golden/resize.npz with the REAL image transform of the reference: the
`transforms.Compose([Resize(input_size), ToTensor()])` built in mmdyn/pytorch/utils/datasets.py:23-31,
applied through the reference's own `VisuoTactileDataset._parse_list_data` (datasets.py:382-392) and
batched with its `seq_collate_fn` (:395-404) — i.e. Pillow (12.2.0 here) + torchvision (0.26) doing
the arithmetic.  Run in the build container only:

    python tests/golden/make_golden_resize.py

Inputs are regenerated from seeds at test time (`frames_for`), only the outputs are stored, as the
uint8 value v of every output element (ToTensor gives v / 255 exactly)."""
import os
import sys

import numpy as np
import torch
import torchvision
from torchvision import transforms

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
CASES = [  # (name, n_frames, in_h, in_w, out_h, out_w)
    ("sim256", 3, 256, 256, 64, 64),   # the simulator's 256x256 renders (dataset default)
    ("half", 2, 128, 128, 64, 64),
    ("ragged", 2, 200, 300, 64, 64),   # non-integer scale factors, non-square
    ("odd", 2, 255, 257, 64, 64),
    ("upscale", 2, 50, 70, 64, 64),    # filterscale clamps to 1: plain bilinear
    ("mixed", 1, 100, 64, 64, 64),     # horizontal pass skipped
    ("same", 1, 64, 64, 64, 64),       # both passes skipped (Image.resize returns a copy)
    ("rect", 1, 256, 256, 32, 48),
]


def frames_for(name, n, h, w):
    """Seeded uint8 frames: uniform noise, plus a smooth ramp with saturated patches in frame 0."""
    rs = np.random.RandomState(sum(map(ord, name)) * 7919 + n)
    f = rs.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    f[0, :, :, 0] = (255 * xx / max(w - 1, 1)).astype(np.uint8)
    f[0, :, :, 1] = (255 * yy / max(h - 1, 1)).astype(np.uint8)
    f[0, : h // 4, : w // 4, :] = 255
    f[0, -(h // 4):, -(w // 4):, :] = 0
    return f


def main():
    sys.path.insert(0, REF)
    from mmdyn.pytorch.utils.datasets import VisuoTactileDataset, seq_collate_fn
    out = {}
    for name, n, h, w, oh, ow in CASES:
        ds = object.__new__(VisuoTactileDataset)
        ds.transform = transforms.Compose([torchvision.transforms.Resize((oh, ow)), transforms.ToTensor()])
        frames = frames_for(name, n, h, w)
        # one "sequence" per frame, data = [image, pose-like vector] as the dataset stores it
        seqs = [([torch.stack([ds._parse_list_data([frames[i], np.zeros(7)])[0]])], [torch.zeros(1)]) for i in range(n)]
        data, _ = seq_collate_fn(seqs)
        t = data[0]
        assert t.shape == (n, 3, oh, ow) and t.dtype == torch.float32
        v = torch.round(t * 255).to(torch.uint8)
        assert torch.equal(v.float() / 255, t)
        out[name] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "resize.npz"), **out)
    print("wrote resize.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
