"""Model factory and the pose `Regressor` baseline — mirror of mmdyn/pytorch/models/models.py
(:9-25 factory, :28-77 Regressor).  As everywhere in mmdyn_b200, the nn.Modules hold the parameters
(same names / shapes / init order as the reference) and the math runs in libmmdyn_b200.so."""
import torch
from torch import nn

from mmdyn_b200 import engine, noise
from mmdyn_b200.pytorch import config
from mmdyn_b200.pytorch.models.vae import VAE, MVAE, Swish, _anchor, _require_cuda  # noqa: F401


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def setup_model(model_name, cross_modal=False, **kwargs):
    """Same dispatch rules as the reference: 'mvae' + cross-modal input -> MVAE, any other
    '...vae' -> VAE (which refuses cross-modal input), 'regressor' -> Regressor."""
    assert (model_name in config.MODELS), "Model is not implement yet"
    if 'mvae' in model_name and cross_modal:
        return MVAE(**kwargs)
    if 'vae' in model_name:
        assert not cross_modal, "VAE does not work with cross modal inputs."
        return VAE(**kwargs)
    if 'regressor' in model_name:
        return Regressor(**kwargs)
    raise SystemExit("The model and modality combination is not valid.")


class _RegressorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, ex, mask, c):
        rec = ex.forward(x.contiguous().float(), mask, engine.FreshAlloc(x.device), "reg", True, c)
        ctx.ex, ctx.rec = ex, rec
        return rec["out"]

    @staticmethod
    def backward(ctx, d_out):
        ex, rec = ctx.ex, ctx.rec
        ex.arena.attach_grads()
        ex.backward(rec, d_out.contiguous().float(), engine.FreshAlloc(d_out.device), "reg", 1.0)
        return None, None, None, None, None


class Regressor(nn.Module):
    """models.py:28-77: DCGAN-style conv trunk (the cnn Encoder's conv_net / fc_net) + a 3-layer MLP.

    `num_classes` is the width of the condition concatenated behind fc_net when conditional=True
    (models.py:36).  The reference's own Regression.set_model passes it as `condition_dim`
    (problems.py:272-277), a keyword Regressor.__init__ does not take, so `--problem-type regression`
    raises TypeError there as shipped; the mirror accepts both spellings."""

    def __init__(self, out_dim=7, conditional=False, num_classes=None, condition_dim=None):
        super().__init__()
        self.conditional = conditional
        self.num_classes = num_classes if num_classes is not None else condition_dim
        if conditional and not 1 <= int(self.num_classes or 0) <= 8:
            raise NotImplementedError("conditional Regressor: 1..8 real-valued condition components (the shock force)")
        cnn_features_out = 256 * 5 * 5
        cnn_features_comp = 512 + self.conditional * (self.num_classes or 0)
        self.conv_net = nn.Sequential(
            nn.Conv2d(3, 32, 4, 2, 1, bias=False),
            Swish(),
            nn.Conv2d(32, 64, 4, 2, 1, bias=False),
            nn.BatchNorm2d(64),
            Swish(),
            nn.Conv2d(64, 128, 4, 2, 1, bias=False),
            nn.BatchNorm2d(128),
            Swish(),
            nn.Conv2d(128, 256, 4, 1, 0, bias=False),
            nn.BatchNorm2d(256),
            Swish()
        )
        self.fc_net = nn.Sequential(
            nn.Linear(cnn_features_out, 512),
            Swish(),
            nn.Dropout(p=0.1),
        )
        self.out_net = nn.Sequential(
            nn.Linear(cnn_features_comp, 256),
            nn.ReLU(),
            nn.Linear(256, 256),
            nn.ReLU(),
            nn.Linear(256, out_dim)
        )
        self.noise = None  # None -> mmdyn_b200.noise.get_default()

    def forward(self, x, c=None):
        _require_cuda(x, "regressor input")
        if not self.training:
            raise NotImplementedError("eval-mode BatchNorm is not part of the reference path "
                                      "(the reference never leaves train mode: problems.py:145,174)")
        arena, ex = engine.get_execs(self, x.device)
        cond = None
        if self.conditional:
            if c is None:
                raise ValueError("conditional=True: a condition tensor is required")
            _require_cuda(c, "condition")
            c = c.unsqueeze(1) if c.dim() == 1 else c
            if tuple(c.shape) != (x.size(0), self.num_classes):
                raise ValueError(f"condition has shape {tuple(c.shape)}, expected ({x.size(0)}, {self.num_classes})")
            cond = c.float().contiguous()
        src = self.noise if self.noise is not None else noise.get_default()
        mask = src.dropout_mask(x.size(0), x.device)
        return _RegressorFn.apply(x, _anchor(arena), ex["reg"], mask, cond)
