"""Dataset plumbing around the step.

The reference's PNG+JSON loader (`mmdyn/pytorch/utils/datasets.py`) is host-side I/O outside the
accelerated path; this module only (a) delegates to it when the reference
package is importable and `dataset_path` is a real directory, and (b) provides a synthetic,
dataset-shaped stand-in (`--dataset-path synthetic[:n_sequences[:seq_length]]`) whose batches obey
the contract of `seq_collate_fn` (datasets.py:395-404): lists of (B*L, ...) tensors

    data   = [visual (B*L,3,64,64), tactile (same), pose (B*L,7), available (B*L,2)]
    target = [visual_final, tactile_final, pose_final (B*L,7), seg_mask (B*L,3,64,64)]
"""
import os

import torch
from torch.utils.data import DataLoader, Dataset


class SyntheticVisuoTactileDataset(Dataset):
    def __init__(self, n_sequences=256, seq_length=50, seed=0):
        self.n, self.L = n_sequences, seq_length
        self.seed = seed
        self.seq_length = seq_length
        # class labels per sequence: Reconstruction._set_condition_dim reads `.targets`
        self.targets = [0] * n_sequences
        self.data = None

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
        target = [vis[-1:].expand(L, -1, -1, -1), tac[-1:].expand(L, -1, -1, -1), pose[-1:].expand(L, -1), mask]
        return data, target


def seq_collate_fn(batch):
    """Concatenate sequences along dim 0 -> (B*L, ...) per field (contract of datasets.py:395-404)."""
    n_data, n_tgt = len(batch[0][0]), len(batch[0][1])
    data = [torch.cat([b[0][k] for b in batch], 0) for k in range(n_data)]
    target = [torch.cat([b[1][k] for b in batch], 0) for k in range(n_tgt)]
    return data, target


def dataset_setup(dataset_path, problem_type, input_size=(64, 64), batchsize=128, shuffle=True):
    path = os.path.expanduser(str(dataset_path))
    if path.startswith("synthetic"):
        parts = path.split(":")
        n_seq = int(parts[1]) if len(parts) > 1 else 4 * batchsize
        L = int(parts[2]) if len(parts) > 2 else 50
        train = SyntheticVisuoTactileDataset(n_seq, L, seed=0)
        test = SyntheticVisuoTactileDataset(max(batchsize, n_seq // 4), L, seed=1)
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
