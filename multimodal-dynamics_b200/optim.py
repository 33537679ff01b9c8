"""Fused optimizers over the flat parameter arena (one kernel launch per step).

FusedAdam / FusedSGD are torch.optim.Optimizer subclasses so that `Problem.set_optimizer`
(problems.py:130-138) and user code keep their shape; the update itself is `mmdyn_adam_flat` /
`mmdyn_sgd_flat`: a single float4-vectorised pass over parameters, gradients and moments
(28 B/parameter for Adam) instead of torch's multi-tensor foreach sequence."""
import torch

from . import engine, ops


class _FlatOptimizer(torch.optim.Optimizer):
    def __init__(self, model, defaults):
        self.model = model
        params = list(model.parameters())
        super().__init__(params, defaults)
        self._step = 0
        self._bufs = None
        self.grad_prescale = 1.0  # e.g. 1/world_size after a sum all-reduce

    def _arena(self):
        arena = engine.get_arena(self.model)
        if self._bufs is None or self._bufs[0].device != arena.flat.device or self._bufs[0].numel() != arena.total:
            self._bufs = tuple(torch.zeros_like(arena.flat) for _ in range(self._n_bufs))
            self._step = 0
            # the step count also lives on the device so that a captured CUDA graph advances it
            self._step_dev = torch.zeros(1, dtype=torch.int64, device=arena.flat.device)
            # sticky device-side flag: set by the update kernel when a gradient entry is inf / NaN
            self._flag = torch.zeros(1, dtype=torch.int32, device=arena.flat.device)
        return arena

    def nonfinite_flag(self):
        """0-dim int32 device tensor: != 0 once any update since the last check_finite() met a non-finite
        gradient entry (that entry was skipped).  Reading it is a host sync — do it where the loss is read."""
        self._arena()
        return self._flag[0]

    def check_finite(self, loss=None):
        """Raise FloatingPointError if a non-finite gradient entry was met since the last call (or if `loss`
        is non-finite).  fp16 operands carry no overflow protection of their own: this is the
        one per-step health signal, and it costs nothing until it is read."""
        self._arena()
        bad = int(self._flag.item()) != 0
        if bad:
            self._flag.zero_()
        lv = float(loss) if loss is not None else 0.0
        if bad or lv != lv or lv in (float("inf"), float("-inf")):
            raise FloatingPointError(
                "mmdyn_b200: non-finite " + ("gradient entries" if bad else "loss") + " in the training step "
                "(fp16 overflow or a poisoned input); the affected parameter entries were NOT updated")

    def state_dict(self):
        """Flat-arena optimizer state (moment buffers + step count) for checkpoint resume."""
        self._arena()
        # the device-side counter is the truth: CUDA-graph replays advance it without touching self._step
        return {"kind": type(self).__name__, "step": int(self._step_dev.item()), "bufs": [b.detach().cpu() for b in self._bufs],
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, state):
        arena = self._arena()
        if state.get("kind") != type(self).__name__ or len(state["bufs"]) != self._n_bufs \
                or state["bufs"][0].numel() != arena.total:
            raise ValueError("optimizer state does not match this optimizer / model")
        for dst, src in zip(self._bufs, state["bufs"]):
            dst.copy_(src)
        self._step = int(state["step"])
        self._step_dev.fill_(self._step)
        for g, sg in zip(self.param_groups, state.get("param_groups", [])):
            g.update(sg)

    def zero_grad(self, set_to_none=True):
        """Keeps p.grad as views of the gradient arena and clears it with one memset."""
        arena = self._arena()
        arena.attach_grads()
        arena.grad.zero_()


class FusedAdam(_FlatOptimizer):
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0) semantics (problems.py:138)."""
    _n_bufs = 2

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(model, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        arena = self._arena()
        arena.attach_grads()
        g = self.param_groups[0]
        self._step += 1
        m, v = self._bufs
        ops.rng_advance(self._step_dev, 1)
        ops.adam_flat_devstep(arena.flat, arena.grad, m, v, arena.total, g["lr"], g["betas"][0], g["betas"][1],
                              g["eps"], g["weight_decay"], self._step_dev, self.grad_prescale, flag=self._flag)
        arena.bump()


class FusedSGD(_FlatOptimizer):
    """torch.optim.SGD(lr, momentum=0.9, weight_decay=5e-4) semantics (problems.py:132-136)."""
    _n_bufs = 1

    def __init__(self, model, lr=1e-3, momentum=0.9, weight_decay=5e-4):
        super().__init__(model, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        arena = self._arena()
        arena.attach_grads()
        g = self.param_groups[0]
        self._step += 1
        # the momentum buffer starts at zero, so momentum*buf + g == g on step 1 (torch's first-step rule)
        ops.sgd_flat(arena.flat, arena.grad, self._bufs[0], arena.total, g["lr"], g["momentum"], g["weight_decay"],
                     False, self.grad_prescale, flag=self._flag)
        arena.bump()
