"""Device-side input pipeline on a real B200: mmdyn_frames_u8_to_f32 against the
oracle (itself pinned to the real PIL + torchvision transform by tests/test_resize_cpu.py) and the
committed golden outputs — bit-exact, including gathered / repeated / empty index sets and ragged sizes;
at dataset scale through size-independent properties."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import resize_oracle as ro  # noqa: E402
from tests.golden.make_golden_resize import CASES, frames_for  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "resize.npz"))
DEV = "cuda"


def _store(frames, size):
    from mmdyn_b200.pytorch.utils.datasets import DeviceFrameStore
    return DeviceFrameStore(frames, size, DEV)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_kernel_matches_reference_goldens(case):
    name, n, h, w, oh, ow = case
    out = _store(frames_for(name, n, h, w), (oh, ow)).images().cpu().numpy()
    assert np.array_equal(out, GOLD[name].astype(np.float32) / np.float32(255.0))


@pytest.mark.parametrize("sizes", [(256, 256, 64, 64), (97, 131, 64, 64), (33, 40, 64, 64), (480, 640, 64, 64),
                                   (64, 64, 64, 64), (30, 17, 30, 17)])  # same size: the ToTensor-only kernel
def test_kernel_matches_oracle_with_gather(sizes):
    h, w, oh, ow = sizes
    rs = np.random.RandomState(h * 1000 + w)
    frames = rs.randint(0, 256, (7, h, w, 3)).astype(np.uint8)
    st = _store(frames, (oh, ow))
    idx = [6, 0, 3, 3, 5]
    out = st.images(idx).cpu().numpy()
    assert np.array_equal(out, ro.frames_to_tensor(frames, oh, ow, index=idx))
    assert st.images([]).shape == (0, 3, oh, ow)           # empty batch
    assert np.array_equal(st.images().cpu().numpy(), ro.frames_to_tensor(frames, oh, ow))


def test_dataset_scale_properties():
    """2048 frames of 256x256 (the dataset's render size): constant frames stay constant (the filter
    weights sum to one in fixed point), the result is invariant to how the batch is split and gathered,
    and a random sample of frames agrees with the oracle bit for bit."""
    n = 2048
    g = torch.Generator(device=DEV).manual_seed(3)
    frames = torch.randint(0, 256, (n, 256, 256, 3), dtype=torch.uint8, device=DEV, generator=g)
    frames[5] = 200
    frames[6] = 0
    frames[7] = 255
    st = _store(frames, (64, 64))
    full = st.images()
    assert torch.all(full[5] == 200.0 / 255.0) and torch.all(full[6] == 0) and torch.all(full[7] == 1.0)
    assert float(full.min()) >= 0.0 and float(full.max()) <= 1.0
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(4))
    parts = torch.cat([st.images(perm[:700]), st.images(perm[700:])])
    assert torch.equal(parts, full[perm.to(DEV)])
    for i in (0, 5, 1023, 2047):
        assert np.array_equal(full[i].cpu().numpy(), ro.frames_to_tensor(frames[i:i + 1].cpu().numpy(), 64, 64)[0])
    # every output byte is an integer multiple of 1/255
    v = full[:64] * 255
    assert torch.equal(v, v.round())


def test_sequence_loader_matches_reference_collate():
    """DeviceSequenceLoader yields what the reference's __getitem__ + seq_collate_fn would: fields in
    order, (B*L, ...) rows sequence-major, images resized + /255, vectors as float32."""
    from mmdyn_b200.pytorch.utils.datasets import DeviceSequenceLoader, synthetic_u8_sequences
    data, targets = synthetic_u8_sequences(6, 3, size=96, seed=11)
    ld = DeviceSequenceLoader(data, targets, batchsize=2, shuffle=False, out_size=(64, 64), device=DEV)
    batches = list(ld)
    assert len(batches) == 3
    for b, (d, t) in enumerate(batches):
        seqs = [2 * b, 2 * b + 1]
        for k in range(4):
            ref = np.stack([data[s][f][k] for s in seqs for f in range(3)])
            got = d[k].cpu().numpy()
            if ref.ndim > 2:
                assert np.array_equal(got, ro.frames_to_tensor(ref, 64, 64))
            else:
                assert np.array_equal(got, ref.astype(np.float32))
        tref = np.stack([targets[s][f][0] for s in seqs for f in range(3)])
        assert np.array_equal(t[0].cpu().numpy(), ro.frames_to_tensor(tref, 64, 64))
        assert t[3].shape == (6, 3, 64, 64)
    # frame_step = L: only the first frame of each sequence (SeqModeling's [::L])
    ld1 = DeviceSequenceLoader(data, targets, batchsize=3, out_size=(64, 64), device=DEV, frame_step=3)
    d, _ = next(iter(ld1))
    assert d[0].shape == (3, 3, 64, 64)
    assert np.array_equal(d[0].cpu().numpy(), ro.frames_to_tensor(np.stack([data[s][0][0] for s in range(3)]), 64, 64))
