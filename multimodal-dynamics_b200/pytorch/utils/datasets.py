"""Dataset plumbing around the step.

The reference's PNG+JSON loader (`mmdyn/pytorch/utils/datasets.py`) is host-side I/O outside the
accelerated path; this module only (a) delegates to it when the reference
package is importable and `dataset_path` is a real directory, and (b) provides a synthetic,
dataset-shaped stand-in (`--dataset-path synthetic[:n_sequences[:seq_length[:shock_dim]]]`) whose batches obey
the contract of `seq_collate_fn` (datasets.py:395-404): lists of (B*L, ...) tensors

    data   = [visual (B*L,3,64,64), tactile (same), pose (B*L,7), available (B*L,2)(, shock (B*L,shock_dim))]
    target = [visual_final, tactile_final, pose_final (B*L,7), seg_mask (B*L,3,64,64)]
"""
import os

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset


class SyntheticVisuoTactileDataset(Dataset):
    def __init__(self, n_sequences=256, seq_length=50, seed=0, shock_dim=0):
        self.n, self.L = n_sequences, seq_length
        self.seed = seed
        self.seq_length = seq_length
        self.shock_dim = int(shock_dim)
        # class labels per sequence: Reconstruction._set_condition_dim reads `.targets`
        self.targets = [0] * n_sequences
        # SeqModeling._set_condition_dim reads len(train_dataset.data[0][0][4]) (problems.py:676-681): the
        # min-max normalised shock force of exp 3 (datasets.py:251-254), one vector per frame
        self.data = [[[None, None, None, None, np.zeros(self.shock_dim, np.float32)]]] if self.shock_dim else None

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        L = self.L
        vis = torch.rand(L, 3, 64, 64, generator=g)
        tac = torch.rand(L, 3, 64, 64, generator=g)
        pose = torch.rand(L, 7, generator=g)
        avail = torch.ones(L, 2)
        mask = (torch.rand(L, 3, 64, 64, generator=g) > 0.5).float()
        data = [vis, tac, pose, avail]
        if self.shock_dim:
            data.append(torch.rand(1, self.shock_dim, generator=g).expand(L, -1).contiguous())  # one shock per sequence
        target = [vis[-1:].expand(L, -1, -1, -1), tac[-1:].expand(L, -1, -1, -1), pose[-1:].expand(L, -1), mask]
        return data, target


def seq_collate_fn(batch):
    """Concatenate sequences along dim 0 -> (B*L, ...) per field (contract of datasets.py:395-404)."""
    n_data, n_tgt = len(batch[0][0]), len(batch[0][1])
    data = [torch.cat([b[0][k] for b in batch], 0) for k in range(n_data)]
    target = [torch.cat([b[1][k] for b in batch], 0) for k in range(n_tgt)]
    return data, target


class DeviceFrameStore:
    """uint8 frames resident in HBM; `images(index)` is the reference's per-frame
    `Compose([Resize(input_size), ToTensor()])` (utils/datasets.py:23-31, 382-392) and the row
    gathering of `seq_collate_fn` (:395-404) as ONE kernel launch (mmdyn_frames_u8_to_f32,
    bit-identical to Pillow's antialiased bilinear + uint8/255), instead of PIL on the host per frame.
    At >100 k samples/s the host path is the bottleneck: a 256x256 render is
    196 KB as uint8 against 49 KB as the fp32 64x64 tensor the model reads."""

    def __init__(self, frames, out_size=(64, 64), device="cuda"):
        """frames: uint8 array / tensor (N, H, W, 3)."""
        from ... import ops
        t = torch.as_tensor(np.ascontiguousarray(frames) if isinstance(frames, np.ndarray) else frames)
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != 3:
            raise ValueError(f"DeviceFrameStore expects uint8 (N, H, W, 3) frames, got {t.dtype} {tuple(t.shape)}")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mmdyn_b200 has no CPU path: DeviceFrameStore needs a CUDA device")
        self.frames = t.contiguous().to(self.device)
        self.out_size = (int(out_size[0]), int(out_size[1]))
        self.table = ops.resize_table(t.shape[1], t.shape[2], *self.out_size).to(self.device)
        self._ops = ops

    def __len__(self):
        return self.frames.shape[0]

    def images(self, index=None, out=None):
        """fp32 (n, 3, h, w) for frames[index] (int64 tensor / list; None = all frames)."""
        if index is not None:
            index = torch.as_tensor(index, dtype=torch.int64).to(self.device)
        n = len(self) if index is None else index.numel()
        if out is None:
            out = torch.empty(n, 3, *self.out_size, dtype=torch.float32, device=self.device)
        if n:
            self._ops.frames_u8_to_f32(self.frames, index, self.table, out)
        return out


class DeviceSequenceLoader:
    """Drop-in for the reference's DataLoader(collate_fn=seq_collate_fn, drop_last=True): iterating
    yields (data, target) lists of (B*L', ...) tensors — already on the GPU.  Built from a dataset in
    the reference's in-memory format (`VisuoTactileDataset.data / .targets`: per sequence a list of
    frames, per frame a list of fields, images as uint8 HxWx3 arrays, everything else 1-D float
    arrays): image fields are packed into DeviceFrameStores, vectors into fp32 device tensors.
    frame_step = L keeps only the first frame of every sequence — what SeqModeling.parse_input's
    `[::L]` selects (problems.py:634-673) — so the unused L-1 frames are never resized."""

    def __init__(self, data, targets, batchsize, shuffle=False, out_size=(64, 64), device="cuda", frame_step=1,
                 seed=0):
        self.B, self.shuffle, self.step = int(batchsize), bool(shuffle), int(frame_step)
        self.n_seq, self.L = len(data), len(data[0])
        self.device = torch.device(device)
        self.gen = torch.Generator().manual_seed(seed)
        self.fields = [self._pack([[fr[k] for fr in seq] for seq in data], out_size) for k in range(len(data[0][0]))]
        self.tfields = [self._pack([[fr[k] for fr in seq] for seq in targets], out_size)
                        for k in range(len(targets[0][0]))]

    def _pack(self, per_seq, out_size):
        first = np.asarray(per_seq[0][0])
        flat = np.stack([np.asarray(x) for seq in per_seq for x in seq])
        if first.ndim > 1:
            return DeviceFrameStore(flat.astype(np.uint8, copy=False), out_size, self.device)
        return torch.from_numpy(flat).float().to(self.device)

    def __len__(self):
        return self.n_seq // self.B

    def _gather(self, f, idx):
        return f.images(idx) if isinstance(f, DeviceFrameStore) else f.index_select(0, idx)

    def __iter__(self):
        order = torch.randperm(self.n_seq, generator=self.gen) if self.shuffle else torch.arange(self.n_seq)
        frames = torch.arange(0, self.L, self.step)
        for b in range(len(self)):
            seqs = order[b * self.B:(b + 1) * self.B]
            idx = (seqs[:, None] * self.L + frames[None, :]).reshape(-1).to(self.device)
            yield [self._gather(f, idx) for f in self.fields], [self._gather(f, idx) for f in self.tfields]


def synthetic_u8_sequences(n_sequences, seq_length, size=256, seed=0):
    """Reference-format in-memory dataset (see DeviceSequenceLoader) of seeded uint8 renders."""
    rs = np.random.RandomState(seed)
    data, targets = [], []
    for _ in range(n_sequences):
        vis = rs.randint(0, 256, (seq_length, size, size, 3)).astype(np.uint8)
        tac = rs.randint(0, 256, (seq_length, size, size, 3)).astype(np.uint8)
        pose = rs.rand(seq_length, 7).astype(np.float32)
        mask = (rs.rand(seq_length, size, size, 1) > 0.5).astype(np.uint8).repeat(3, -1) * 255
        data.append([[vis[i], tac[i], pose[i], np.ones(2, np.float32)] for i in range(seq_length)])
        targets.append([[vis[-1], tac[-1], pose[-1], mask[i]] for i in range(seq_length)])
    return data, targets


def dataset_setup(dataset_path, problem_type, input_size=(64, 64), batchsize=128, shuffle=True):
    path = os.path.expanduser(str(dataset_path))
    if path.startswith("synthetic-u8"):
        # uint8 renders resident on the GPU, resized / normalised / collated by mmdyn_frames_u8_to_f32
        parts = path.split(":")
        n_seq = int(parts[1]) if len(parts) > 1 else 2 * batchsize
        L = int(parts[2]) if len(parts) > 2 else 4
        size = int(parts[3]) if len(parts) > 3 else 256
        tr_d, tr_t = synthetic_u8_sequences(n_seq, L, size, seed=0)
        te_d, te_t = synthetic_u8_sequences(max(batchsize, n_seq // 4), L, size, seed=1)
        train = SyntheticVisuoTactileDataset(n_seq, L, seed=0)
        test = SyntheticVisuoTactileDataset(max(batchsize, n_seq // 4), L, seed=1)
        return {"train_dataset": train, "test_dataset": test,
                "train_loader": DeviceSequenceLoader(tr_d, tr_t, batchsize, shuffle, input_size),
                "test_loader": DeviceSequenceLoader(te_d, te_t, batchsize, False, input_size), "seq_length": L}
    if path.startswith("synthetic"):
        parts = path.split(":")
        n_seq = int(parts[1]) if len(parts) > 1 else 4 * batchsize
        L = int(parts[2]) if len(parts) > 2 else 50
        sd = int(parts[3]) if len(parts) > 3 else 0
        train = SyntheticVisuoTactileDataset(n_seq, L, seed=0, shock_dim=sd)
        test = SyntheticVisuoTactileDataset(max(batchsize, n_seq // 4), L, seed=1, shock_dim=sd)
        # NB: the reference only seq-collates when 'seq' is in the problem type and therefore crashes
        # for dyn_modeling; both problem types need (B*L, ...) batches.
        kw = dict(batch_size=batchsize, collate_fn=seq_collate_fn, drop_last=True, num_workers=0)
        return {"train_dataset": train, "test_dataset": test,
                "train_loader": DataLoader(train, shuffle=shuffle, **kw),
                "test_loader": DataLoader(test, shuffle=False, **kw), "seq_length": L}
    try:
        from mmdyn.pytorch.utils.datasets import dataset_setup as ref_setup  # the reference's own loader
    except Exception as e:  # pragma: no cover - depends on the user's environment
        raise RuntimeError("real datasets are read by the reference's loader (mmdyn.pytorch.utils.datasets); "
                           "install the reference next to mmdyn_b200 or use --dataset-path synthetic") from e
    return ref_setup(dataset_path, problem_type, input_size=input_size, batchsize=batchsize, shuffle=shuffle)
