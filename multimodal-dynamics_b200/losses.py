"""Stand-alone loss terms on the library's kernels, for callers that hold reconstructions and
posteriors already: `Reconstruction._elbo_loss` / `_mvae_elbo_loss` (problems.py:401-458) called
outside the fused step, and the per-sample scoring variant (reduce=False, :415-417, :451-456).

Each term is a torch.autograd.Function around one kernel launch; the only torch arithmetic here is
the chain-rule multiplication by the upstream scalar gradient."""
import torch

from . import ops

F32 = torch.float32


def _flat(t):
    return t.contiguous().float()


class _BCESum(torch.autograd.Function):
    """F.binary_cross_entropy_with_logits(x*m, t*m, reduction='sum') (problems.py:409-413, 445-449)."""

    @staticmethod
    def forward(ctx, logits, target, mask):
        x, t = _flat(logits), _flat(target)
        m = _flat(mask.expand_as(logits)) if mask is not None else None
        n = x.shape[0]
        out = torch.zeros(1, dtype=F32, device=x.device)
        grad = torch.empty_like(x) if logits.requires_grad else None
        ops.bce_logits_flat(x, t, m, out, None, grad, 1.0, n, x.numel() // n)
        ctx.grad, ctx.shape = grad, logits.shape
        return out[0]

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g).view(ctx.shape) if ctx.grad is not None else None, None, None


class _MSESum(torch.autograd.Function):
    """F.mse_loss(r, t, reduction='sum') (problems.py:441-449)."""

    @staticmethod
    def forward(ctx, recon, target):
        r, t = _flat(recon), _flat(target)
        out = torch.zeros(1, dtype=F32, device=r.device)
        grad = torch.empty_like(r) if recon.requires_grad else None
        ops.mse(r, t, out, grad, 1.0, 1.0, r.numel())
        ctx.grad, ctx.shape = grad, recon.shape
        return out[0]

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g).view(ctx.shape) if ctx.grad is not None else None, None


class _KL(torch.autograd.Function):
    """-0.5 * sum(1 + logvar - mu^2 - exp(logvar)) (problems.py:406, 429)."""

    @staticmethod
    def forward(ctx, mu, lv):
        m, l = _flat(mu), _flat(lv)
        B, D = m.shape
        zeros = torch.zeros(B, D, device=m.device)
        o1, o2, o3 = (torch.empty(B, D, device=m.device) for _ in range(3))
        kl = torch.zeros(1, dtype=F32, device=m.device)
        ops.poe_fwd([m], [l], False, D, zeros, o1, o2, o3, None, None, kl, B, D)
        ctx.save_for_backward(m, l, zeros)
        return kl[0]

    @staticmethod
    def backward(ctx, g):
        m, l, zeros = ctx.saved_tensors
        B, D = m.shape
        dm, dl = torch.empty_like(m), torch.empty_like(l)
        ops.poe_bwd([m], [l], False, D, zeros, [None], 1.0, [dm], [dl], D, False, B, D)
        return dm * g, dl * g


def bce_with_logits_sum(logits, target, mask=None):
    return _BCESum.apply(logits, target, mask)


def mse_sum(recon, target):
    return _MSESum.apply(recon, target)


def kl_divergence(mu, logvar):
    return _KL.apply(mu, logvar)


@torch.no_grad()
def bce_with_logits_per_sample(logits, target, mask=None):
    """sum over (1,2,3) of the element-wise BCE: the reduce=False scoring path."""
    x, t = _flat(logits), _flat(target)
    m = _flat(mask.expand_as(logits)) if mask is not None else None
    n = x.shape[0]
    tot = torch.zeros(1, dtype=F32, device=x.device)
    per = torch.zeros(n, dtype=F32, device=x.device)
    ops.bce_logits_flat(x, t, m, tot, per, None, 1.0, n, x.numel() // n)
    return per


@torch.no_grad()
def mse_per_sample(recon, target, mult=1.0):
    r, t = _flat(recon), _flat(target)
    n, d = r.shape
    per = torch.zeros(n, dtype=F32, device=r.device)
    ops.mse_rows(r, t, per, float(mult), n, d)
    return per
