"""Sources of the step's random inputs: Dropout masks (vae.py:213) and the reparametrisation
noise (vae.py:58).

HostNoise   draws from torch's CPU default generator in exactly the reference's order and ships the
            result to the device — what the reference itself does for eps on every device, and
            bit-identical to the CPU reference for the masks (F.dropout(p) on CPU equals
            empty.bernoulli_(1-p)/(1-p) under the same generator state).  Used for parity.
DeviceNoise Philox4x32-10 kernels of libmmdyn_b200.so with the running counter in device memory
            (so a captured CUDA graph draws fresh numbers at every replay).  Used for throughput.
"""
import torch

from . import ops

DROPOUT_P = 0.1


class HostNoise:
    def __init__(self, generator=None):
        self.generator = generator

    def dropout_mask(self, B, device, out=None):
        m = torch.empty(B, 512).bernoulli_(1 - DROPOUT_P, generator=self.generator) / (1 - DROPOUT_P)
        return self._ship(m, device, out)

    def normal(self, B, D, device, out=None):
        return self._ship(torch.randn([B, D], generator=self.generator), device, out)

    @staticmethod
    def _ship(t, device, out):
        if out is None:
            return t.to(device)
        out.copy_(t, non_blocking=False)
        return out


class DeviceNoise:
    def __init__(self, seed=0, device=None):
        self.seed = int(seed)
        self.device = device
        self.ctr = None

    def _counter(self, device):
        if self.ctr is None or self.ctr.device != torch.device(device):
            self.ctr = torch.zeros(1, dtype=torch.int64, device=device)
        return self.ctr

    def state_dict(self):
        """Seed + Philox counter (checkpoint resume: the stream continues where it stopped)."""
        return {"seed": self.seed, "counter": int(self.ctr.item()) if self.ctr is not None else 0}

    @classmethod
    def from_state_dict(cls, state, device):
        src = cls(seed=int(state["seed"]), device=device)
        src._counter(device).fill_(int(state["counter"]))
        return src

    def dropout_mask(self, B, device, out=None):
        out = torch.empty(B, 512, device=device) if out is None else out
        ctr = self._counter(out.device)
        ops.fill_dropout_mask(out, out.numel(), DROPOUT_P, self.seed, 0, ctr)
        ops.rng_advance(ctr, (out.numel() + 3) // 4)
        return out

    def normal(self, B, D, device, out=None):
        out = torch.empty(B, D, device=device) if out is None else out
        ctr = self._counter(out.device)
        ops.fill_normal(out, out.numel(), self.seed, 0, ctr)
        ops.rng_advance(ctr, (out.numel() + 3) // 4)
        return out


    def fill_step(self, mask_bufs, eps):
        """All dropout masks and reparametrisation noise of one step: one launch per buffer."""
        ctr = self._counter(eps.device)
        off = 0
        for mb in mask_bufs:
            ops.fill_dropout_mask(mb, mb.numel(), DROPOUT_P, self.seed, off, ctr)
            off += (mb.numel() + 3) // 4
        ops.fill_normal(eps, eps.numel(), self.seed, off, ctr)
        off += (eps.numel() + 3) // 4
        ops.rng_advance(ctr, off)


_default = HostNoise()


def get_default():
    return _default


def set_default(src):
    global _default
    _default = src
    return src
