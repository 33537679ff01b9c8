"""Variational Autoencoder (VAE) and Multimodal VAE (MVAE) — B200-native mirror of the reference's
`mmdyn/pytorch/models/vae.py` (same class names, constructor kwargs, forward / inference
signatures, parameter order and `state_dict()` keys).

The nn.Conv2d / nn.BatchNorm2d / nn.Linear / nn.ConvTranspose2d objects below are *parameter
holders only* (so initialisation, `.to()`, `state_dict()` and optimizers behave exactly like the
reference); the forward and backward math runs in libmmdyn_b200.so through
`mmdyn_b200.engine` — there is no PyTorch compute fallback and no CPU path.

Scope: architecture 'cnn' for the image experts, the 'mlp' pose expert of MVAE,
models, with or without CVAE shock conditioning (`conditional=True`, vae.py:231-237, 286-291: the
condition columns of the two head Linears and of the decoder's upsample Linear are a rank-cd fp32
term next to the tensor-core GEMM).  A standalone 'mlp' VAE and categorical (one-hot) conditions are
reference features outside this path and raise NotImplementedError.
"""
import torch
import torch.nn as nn

from mmdyn_b200 import engine, noise, ops, plan
from mmdyn_b200.pytorch import config

F16, F32 = torch.float16, torch.float32


def mlp(sizes, activation, output_activation=nn.Identity):
    """vae.py:14-19 (holder layout of the pose expert)."""
    layers = []
    for j in range(len(sizes) - 1):
        act = activation if j < len(sizes) - 2 else output_activation
        layers += [nn.Linear(sizes[j], sizes[j + 1]), act()]
    return nn.Sequential(*layers)


def count_vars(module):
    return sum(p.numel() for p in module.parameters())


class Swish(nn.Module):
    """https://arxiv.org/abs/1710.05941 (vae.py:331-334).  Inside Encoder / Decoder it is fused into
    the BatchNorm-apply kernel; the standalone call runs the same kernel on an fp16 copy."""

    def forward(self, x):
        return _SwishFn.apply(x)


class _SwishFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        xf = x.contiguous().float()
        n = xf.numel()
        pad = (-n) % 8
        xh = torch.zeros(n + pad, dtype=F16, device=x.device)
        ops.f32_to_f16(xf, xh, n)
        yh = torch.empty_like(xh)
        ops.bn_swish_fwd(xh, None, yh, 1, (n + pad) // 8, 8)
        ctx.save_for_backward(xh)
        ctx.n, ctx.shape = n, x.shape
        return yh[:n].float().view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        (xh,) = ctx.saved_tensors
        n = ctx.n
        dh = torch.zeros_like(xh)
        ops.f32_to_f16(dy.contiguous().float(), dh, n)
        ops.bn_swish_bwd_reduce(xh, None, None, dh, None, 1, xh.numel() // 8, 8)
        return dh[:n].float().view(ctx.shape)


def _root_of(mod):
    r = mod.__dict__.get("_mmdyn_root")
    r = r() if r is not None else None
    return (r, mod.__dict__["_mmdyn_prefix"]) if r is not None else (mod, "")


def _bind_children(root):
    import weakref
    for name, child in root.named_children():
        if isinstance(child, (Encoder, Decoder)):
            child.__dict__["_mmdyn_root"] = weakref.ref(root)
            child.__dict__["_mmdyn_prefix"] = name


def _anchor(arena):
    return arena.params[0]


def _condition(mod, c, n):
    """The reference's handling of `c` in Encoder/Decoder.forward (vae.py:231-237, 286-291) for
    real-valued conditions: (n,) -> (n, 1), float, concatenated after the features."""
    if not mod.conditional:
        return None
    if mod.categorical_conditions:
        raise NotImplementedError("categorical (one-hot) conditions are not part of the resting-state path "
                                  "(SeqModeling / DynModeling set categorical_conditions=False: problems.py:677)")
    if c is None:
        raise ValueError("conditional=True: a condition tensor is required")
    _require_cuda(c, "condition")
    c = c.unsqueeze(1) if c.dim() == 1 else c
    if tuple(c.shape) != (n, mod.condition_dim):
        raise ValueError(f"condition has shape {tuple(c.shape)}, expected ({n}, {mod.condition_dim})")
    return c.float().contiguous()


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"mmdyn_b200: {what} must live on a CUDA device (B200); there is no CPU path")


# ---------------------------------------------------------------------------------------------
# autograd bridges (module-level API): activations are owned by the autograd ctx
# ---------------------------------------------------------------------------------------------
class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, ex, mask, track, c=None):
        rec = ex.forward(x.contiguous().float(), [mask], engine.FreshAlloc(x.device), "enc", track, c)
        ctx.ex, ctx.rec = ex, rec
        return rec["heads"]

    @staticmethod
    def backward(ctx, d_heads):
        ex, rec = ctx.ex, ctx.rec
        ex.arena.attach_grads()
        gs = float(rec["B"])
        ex.backward(rec, d_heads.contiguous(), engine.FreshAlloc(d_heads.device), "enc", 1.0 / gs, gs)
        return None, None, None, None, None, None


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, anchor, ex, track, c=None):
        B = z.shape[0]
        zf = z.contiguous().float()
        zh = torch.empty(B, z.shape[1], dtype=F16, device=z.device)
        ops.f32_to_f16(zf, zh, zf.numel())
        rec = ex.forward(zh, 1, B, engine.FreshAlloc(z.device), "dec", track, cond=c)
        ctx.ex, ctx.rec = ex, rec
        return rec["logits"]

    @staticmethod
    def backward(ctx, dlogits):
        ex, rec = ctx.ex, ctx.rec
        ex.arena.attach_grads()
        B = rec["B"]
        gs = float(B)
        dl8 = torch.zeros(B, 66, 66, plan.LOGIT_CP, dtype=F16, device=dlogits.device)  # zero border: engine.DecoderExec.backward
        ops.logit_grad_pack(dlogits.contiguous().float(), dl8, gs, B, 64, 64, 1, cp=plan.LOGIT_CP)
        dz = ex.backward(rec, dl8, engine.FreshAlloc(dlogits.device), "dec", 1.0 / gs)
        ops.scale_f32(dz, dz.numel(), 1.0 / gs)
        return dz, None, None, None, None


class _PoseEncFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, anchor, pex):
        rec = pex.enc_forward(pose.contiguous().float(), engine.FreshAlloc(pose.device), "penc")
        ctx.pex, ctx.rec = pex, rec
        return rec["heads"]

    @staticmethod
    def backward(ctx, d_heads):
        ctx.pex.arena.attach_grads()
        ctx.pex.enc_backward(ctx.rec, d_heads.contiguous(), engine.FreshAlloc(d_heads.device), "penc", 1.0)
        return None, None, None


class _PoseDecFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, anchor, pex):
        rec = pex.dec_forward(z.contiguous().float(), engine.FreshAlloc(z.device), "pdec")
        ctx.pex, ctx.rec = pex, rec
        return rec["rec"]

    @staticmethod
    def backward(ctx, d_rec):
        ctx.pex.arena.attach_grads()
        dz = ctx.pex.dec_backward(ctx.rec, d_rec.contiguous(), engine.FreshAlloc(d_rec.device), "pdec", 1.0)
        return dz, None, None


class _PoEReparamFn(torch.autograd.Function):
    """(mu, logvar, z) from the experts' [mu | logvar] rows (vae.py:139-159, 311-318)."""

    @staticmethod
    def forward(ctx, use_prior, eps, D, *experts):
        # experts: tensors (B, 2*D) fp32 row-major = [mu | logvar]
        B = experts[0].shape[0] if experts else eps.shape[0]
        dev = eps.device
        ex = [e.contiguous() for e in experts]
        mu, lv, z = (torch.empty(B, D, device=dev) for _ in range(3))
        kl = torch.zeros(1, device=dev)
        ops.poe_fwd([e[:, :D] for e in ex], [e[:, D:] for e in ex], use_prior, 2 * D, eps, mu, lv, z, None, None,
                    kl, B, D)
        ctx.use_prior, ctx.D, ctx.ex, ctx.eps = use_prior, D, ex, eps
        return mu, lv, z

    @staticmethod
    def backward(ctx, dmu, dlv, dz):
        ex, D = ctx.ex, ctx.D
        B = ctx.eps.shape[0]
        grads = [torch.empty_like(e) for e in ex]
        if ex:
            ops.poe_bwd([e[:, :D] for e in ex], [e[:, D:] for e in ex], ctx.use_prior, 2 * D, ctx.eps,
                        [dz.contiguous() if dz is not None else None], 0.0, [g[:, :D] for g in grads],
                        [g[:, D:] for g in grads], 2 * D, False, B, D,
                        dmu_in=dmu.contiguous() if dmu is not None else None,
                        dlv_in=dlv.contiguous() if dlv is not None else None)
        return (None, None, None) + tuple(grads)


# ---------------------------------------------------------------------------------------------
# modules
# ---------------------------------------------------------------------------------------------
class Autoencoder(nn.Module):
    """Base class for Autoencoders (vae.py:26-67)."""

    def __init__(self, input_dim=784, encoder_hid=[256, 256], latent_size=8,
                 decoder_hid=[256, 256], condition_dim=None, architecture='mlp',
                 conditional=False, categorical_conditions=False):
        super().__init__()
        assert type(encoder_hid) == list
        assert type(latent_size) == int
        assert type(decoder_hid) == list
        assert architecture in config.ARCHITECTURES
        self.latent_size = latent_size
        self.input_dim = input_dim
        self.condition_dim = condition_dim
        self.architecture = architecture
        self.conditional = conditional
        self.categorical_conditions = categorical_conditions
        self.noise = None  # None -> mmdyn_b200.noise.get_default()

    def _noise(self):
        return self.noise if self.noise is not None else noise.get_default()

    def reparametrize(self, means, log_var):
        """vae.py:52-61: z = eps * exp(0.5 * log_var) + means, eps ~ N(0, I)."""
        _require_cuda(means, "means")
        eps = self._noise().normal(means.size(0), self.latent_size, means.device)
        heads = torch.cat((means, log_var), dim=1)
        _, _, z = _PoEReparamFn.apply(False, eps, self.latent_size, heads)
        return z

    def forward(self, x):
        raise NotImplementedError

    def inference(self, n=1):
        raise NotImplementedError


class VAE(Autoencoder):
    """Vanilla VAE (vae.py:70-98)."""

    def __init__(self, use_pose=False, **kwargs):
        super().__init__(**kwargs)
        if kwargs.get('architecture', 'mlp') != 'cnn':
            raise NotImplementedError("mmdyn_b200 accelerates the cnn-vae / cnn-mvae path only")
        self.encoder = Encoder(**kwargs)
        self.decoder = Decoder(**kwargs)
        _bind_children(self)

    def forward(self, x, c=None):
        means, log_var = self.encoder(x, c)
        z = self.reparametrize(means, log_var)
        recon_x = self.decoder(z, c)
        return recon_x, means, log_var

    def inference(self, n=1, c=None):
        dev = next(self.parameters()).device
        z = self._noise().normal(n, self.latent_size, dev)
        return self.decoder(z, c)


class MVAE(Autoencoder):
    """'Multimodal Generative Models for Scalable Weakly-Supervised Learning' (vae.py:101-176)."""

    def __init__(self, use_pose=False, **kwargs):
        super().__init__(**kwargs)
        assert kwargs['architecture'] != 'mlp', "MVAE is not implemented with MLP"
        self._use_pose = use_pose
        self.visual_encoder = Encoder(**kwargs)
        self.visual_decoder = Decoder(**kwargs)
        self.tactile_encoder = Encoder(**kwargs)
        self.tactile_decoder = Decoder(**kwargs)
        if self._use_pose:
            self.pose_encoder = Encoder(input_dim=7, layer_sizes=[512, 512],
                                        latent_size=kwargs["latent_size"],
                                        condition_dim=0, architecture="mlp")
            self.pose_decoder = Decoder(output_dim=7, layer_sizes=[512, 512],
                                        latent_size=kwargs["latent_size"],
                                        condition_dim=0, architecture="mlp")
        self.experts = ProductOfExperts()
        _bind_children(self)

    def forward(self, x, pose=None, condition=None):
        assert isinstance(x, list) or isinstance(x, tuple)
        visual, tactile = x
        if visual is not None:
            batch_size, dev = visual.size(0), visual.device
        elif tactile is not None:
            batch_size, dev = tactile.size(0), tactile.device
        else:
            batch_size, dev = pose.size(0), pose.device
        # experts in the reference's order: (implicit) prior, visual, tactile, pose (vae.py:139-154)
        heads = []
        if visual is not None:
            heads.append(self.visual_encoder._heads(visual, condition))
        if tactile is not None:
            heads.append(self.tactile_encoder._heads(tactile, condition))
        if pose is not None and self._use_pose:
            heads.append(self.pose_encoder._heads(pose))
        eps = self._noise().normal(batch_size, self.latent_size, dev)
        means, log_var, z = _PoEReparamFn.apply(True, eps, self.latent_size, *heads)
        visual_recon = self.visual_decoder(z, c=condition)
        tactile_recon = self.tactile_decoder(z, c=condition)
        pose_recon = self.pose_decoder(z, c=condition) if self._use_pose else None
        return visual_recon, tactile_recon, pose_recon, means, log_var

    def inference(self, n=1, c=None):
        dev = next(self.parameters()).device
        z = self._noise().normal(n, self.latent_size, dev)
        return self.visual_decoder(z, c), self.tactile_decoder(z, c)


class Encoder(nn.Module):
    """vae.py:179-242."""

    def __init__(self, input_dim=784, layer_sizes=[256, 256], latent_size=8,
                 architecture='mlp', conditional=False, categorical_conditions=False,
                 condition_dim=None, **kwargs):
        super().__init__()
        self.architecture = architecture
        self.conditional = conditional
        self.categorical_conditions = categorical_conditions
        self.condition_dim = condition_dim
        self.latent_size = latent_size
        if conditional and (architecture != 'cnn' or categorical_conditions or not condition_dim
                            or not 1 <= int(condition_dim) <= 8):
            raise NotImplementedError("conditional encoders: cnn architecture with 1..8 real-valued condition "
                                      "components (the shock force, problems.py:676-681) only")
        if architecture == 'cnn':
            if latent_size != 256:
                raise NotImplementedError("the sm_100a kernels are specialised for latent_size=256 (main.py:49)")
            cnn_features_out = 256 * 5 * 5
            cnn_features_comp = 512 + self.conditional * (self.condition_dim or 0)
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
            self.linear_means = nn.Linear(cnn_features_comp, latent_size)
            self.linear_log_var = nn.Linear(cnn_features_comp, latent_size)
        else:
            if input_dim != 7 or list(layer_sizes) != [512, 512] or latent_size != 256:
                raise NotImplementedError("only the pose expert MLP (7 -> 512 -> 512 -> 256) is accelerated")
            layer_sizes = [input_dim] + layer_sizes
            self.fc_net = mlp(layer_sizes, nn.ReLU, nn.Identity)
            self.linear_means = nn.Linear(layer_sizes[-1], latent_size)
            self.linear_log_var = nn.Linear(layer_sizes[-1], latent_size)

    def _heads(self, x, c=None):
        _require_cuda(x, "encoder input")
        root, prefix = _root_of(self)
        arena, ex = engine.get_execs(root, x.device)
        if self.architecture == 'cnn':
            if not self.training:
                raise NotImplementedError("eval-mode BatchNorm is not part of the reference path "
                                          "(the reference never leaves train mode: problems.py:145,174)")
            src = root.noise if getattr(root, "noise", None) is not None else noise.get_default()
            mask = src.dropout_mask(x.size(0), x.device)
            return _EncoderFn.apply(x, _anchor(arena), ex["enc"][prefix or "encoder"], mask, True,
                                    _condition(self, c, x.size(0)))
        return _PoseEncFn.apply(x, _anchor(arena), ex["pose"])

    def forward(self, x, c=None):
        h = self._heads(x, c)
        return h[:, :self.latent_size], h[:, self.latent_size:]


class Decoder(nn.Module):
    """vae.py:245-301.  Returns logits (the Sigmoid is commented out in the reference, :278)."""

    def __init__(self, output_dim=784, layer_sizes=[256, 256], latent_size=2,
                 architecture='mlp', conditional=False, categorical_conditions=False,
                 condition_dim=None, **kwargs):
        super().__init__()
        self.architecture = architecture
        self.conditional = conditional
        self.categorical_conditions = categorical_conditions
        self.condition_dim = condition_dim
        if conditional and (architecture != 'cnn' or categorical_conditions or not condition_dim
                            or not 1 <= int(condition_dim) <= 8):
            raise NotImplementedError("conditional decoders: cnn architecture with 1..8 real-valued condition "
                                      "components (the shock force, problems.py:676-681) only")
        if architecture == 'cnn':
            if latent_size != 256:
                raise NotImplementedError("the sm_100a kernels are specialised for latent_size=256 (main.py:49)")
            self.upsample = nn.Sequential(
                nn.Linear(latent_size + self.conditional * (self.condition_dim or 0), 256 * 5 * 5),
                Swish()
            )
            self.hallucinate = nn.Sequential(
                nn.ConvTranspose2d(256, 128, 4, 1, 0, bias=False),
                nn.BatchNorm2d(128),
                Swish(),
                nn.ConvTranspose2d(128, 64, 4, 2, 1, bias=False),
                nn.BatchNorm2d(64),
                Swish(),
                nn.ConvTranspose2d(64, 32, 4, 2, 1, bias=False),
                nn.BatchNorm2d(32),
                Swish(),
                nn.ConvTranspose2d(32, 3, 4, 2, 1, bias=False),
                # nn.Sigmoid()
            )
        else:
            if output_dim != 7 or list(layer_sizes) != [512, 512] or latent_size != 256:
                raise NotImplementedError("only the pose expert MLP (256 -> 512 -> 512 -> 7) is accelerated")
            layer_sizes = [latent_size] + layer_sizes + [output_dim]
            self.deconv_net = mlp(layer_sizes, nn.ReLU, nn.Identity)

    def forward(self, z, c=None):
        _require_cuda(z, "decoder input")
        root, prefix = _root_of(self)
        arena, ex = engine.get_execs(root, z.device)
        if self.architecture == 'cnn':
            if not self.training:
                raise NotImplementedError("eval-mode BatchNorm is not part of the reference path "
                                          "(the reference never leaves train mode: problems.py:145,174)")
            return _DecoderFn.apply(z, _anchor(arena), ex["dec"][prefix or "decoder"], True,
                                    _condition(self, c, z.size(0)))
        return _PoseDecFn.apply(z, _anchor(arena), ex["pose"])


class ProductOfExperts(nn.Module):
    """Parameters of a product of independent Gaussian experts (vae.py:304-318).
    mu, logvar: M x B x D for 2 <= M <= 4 experts (the prior is just another expert here)."""

    def forward(self, mu, logvar, eps=1e-8):
        if eps != 1e-8:
            raise NotImplementedError("the fused kernel implements the reference's eps=1e-8")
        M, B, D = mu.shape
        if not 2 <= M <= 4:
            raise NotImplementedError("ProductOfExperts kernel handles 2..4 stacked experts")
        _require_cuda(mu, "experts")
        heads = [torch.cat((mu[m], logvar[m]), dim=1) for m in range(M)]
        zeros = torch.zeros(B, D, device=mu.device)
        pd_mu, pd_logvar, _ = _PoEReparamFn.apply(False, zeros, D, *heads)
        return pd_mu, pd_logvar


def prior_expert(size, device=torch.device('cpu')):
    """Universal prior expert N(0, 1) (vae.py:321-328).  MVAE.forward does not materialise it (the
    fused kernel adds its precision term implicitly); kept for API compatibility."""
    mu = torch.zeros(size).to(device)
    logvar = torch.zeros(size).to(device)
    return mu, logvar
