"""ORACLE — test infrastructure, not product code.

A CPU restatement (torch fp32/fp64 functional ops, no nn.Module, no autograd graph tricks) of the
cnn-vae / cnn-mvae training + inference step of SAIC-MONTREAL/multimodal-dynamics.  Every function
cites the reference file:line it follows (paths relative to the reference root).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module, and only as the checker or the timed CPU baseline — the product path
(`multimodal-dynamics_b200/`) never does, and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors for `mmdyn/pytorch`, so
this oracle is pinned against the reference ITSELF, imported live from /root/reference in the build
container by `tests/golden/make_golden.py`; the resulting fixtures are committed under
`tests/golden/` and checked by `tests/test_oracle_cpu.py` wherever /root/reference is absent.

State is a plain dict of tensors with the reference's `state_dict()` key names, e.g.
`visual_encoder.conv_net.0.weight`, `pose_decoder.deconv_net.4.bias`.
"""
import math

import torch
import torch.nn.functional as F

POE_EPS = 1e-8  # vae.py:311
BN_EPS = 1e-5   # nn.BatchNorm2d default (vae.py:201)
BN_MOMENTUM = 0.1
DROPOUT_P = 0.1  # vae.py:213


def swish(x):
    """vae.py:331-334."""
    return x * torch.sigmoid(x)


def _bn_train(x, sd, prefix, track):
    """nn.BatchNorm2d in training mode (vae.py:201,204,207,269,272,275): biased variance for the
    normalisation, unbiased for running_var, momentum 0.1; updates running stats in `sd` in place
    when `track` (the reference is never put in eval mode: problems.py:145,174)."""
    rm = sd[prefix + ".running_mean"] if track else None
    rv = sd[prefix + ".running_var"] if track else None
    y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training=True,
                     momentum=BN_MOMENTUM, eps=BN_EPS)
    if track:
        sd[prefix + ".num_batches_tracked"] += 1
    return y


def _cat_condition(h, c):
    """Real-valued condition handling of Encoder / Decoder.forward (vae.py:231-237, 286-291):
    (n,) -> (n,1), then torch.cat((features, c.float()), -1).  c = None: un-conditional model."""
    if c is None:
        return h
    c = c.unsqueeze(1) if c.dim() == 1 else c
    return torch.cat((h, c.to(h.dtype)), dim=-1)


def encoder_cnn(sd, p, x, dropout_mask=None, track=True, acts=None, c=None):
    """Encoder.forward, architecture 'cnn' (vae.py:197-216, 224-242).  dropout_mask: (B,512) tensor
    with values 0 or 1/(1-p) (None = no dropout); c: condition rows for conditional=True models
    (heads are Linear(512 + condition_dim, 256), vae.py:196); returns (means, log_vars)."""
    def rec(k, v):
        if acts is not None:
            acts[p + "." + k] = v
    h = F.conv2d(x, sd[p + ".conv_net.0.weight"], stride=2, padding=1)
    rec("conv1", h)
    h = swish(h)
    h = F.conv2d(h, sd[p + ".conv_net.2.weight"], stride=2, padding=1)
    rec("conv2", h)
    h = swish(_bn_train(h, sd, p + ".conv_net.3", track))
    rec("act2", h)
    h = F.conv2d(h, sd[p + ".conv_net.5.weight"], stride=2, padding=1)
    rec("conv3", h)
    h = swish(_bn_train(h, sd, p + ".conv_net.6", track))
    rec("act3", h)
    h = F.conv2d(h, sd[p + ".conv_net.8.weight"], stride=1, padding=0)
    rec("conv4", h)
    h = swish(_bn_train(h, sd, p + ".conv_net.9", track))
    rec("act4", h)
    h = h.reshape(h.size(0), -1)
    h = swish(F.linear(h, sd[p + ".fc_net.0.weight"], sd[p + ".fc_net.0.bias"]))
    rec("fc", h)
    if dropout_mask is not None:
        h = h * dropout_mask
    h = _cat_condition(h, c)
    mu = F.linear(h, sd[p + ".linear_means.weight"], sd[p + ".linear_means.bias"])
    lv = F.linear(h, sd[p + ".linear_log_var.weight"], sd[p + ".linear_log_var.bias"])
    rec("mu", mu)
    rec("lv", lv)
    return mu, lv


def encoder_mlp(sd, p, x):
    """pose Encoder, architecture 'mlp' with layer_sizes [7,512,512] (vae.py:14-19, 118-120,
    219-222): Linear+ReLU, Linear+Identity, then the two heads."""
    h = F.relu(F.linear(x, sd[p + ".fc_net.0.weight"], sd[p + ".fc_net.0.bias"]))
    h = F.linear(h, sd[p + ".fc_net.2.weight"], sd[p + ".fc_net.2.bias"])
    mu = F.linear(h, sd[p + ".linear_means.weight"], sd[p + ".linear_means.bias"])
    lv = F.linear(h, sd[p + ".linear_log_var.weight"], sd[p + ".linear_log_var.bias"])
    return mu, lv


def decoder_cnn(sd, p, z, track=True, acts=None, c=None):
    """Decoder.forward, architecture 'cnn' (vae.py:263-279, 286-296): returns LOGITS (no sigmoid).
    c: condition rows for conditional=True models (upsample is Linear(256 + condition_dim, 6400), :257)."""
    def rec(k, v):
        if acts is not None:
            acts[p + "." + k] = v
    z = _cat_condition(z, c)
    h = swish(F.linear(z, sd[p + ".upsample.0.weight"], sd[p + ".upsample.0.bias"]))
    h = h.view(-1, 256, 5, 5)
    rec("up", h)
    h = F.conv_transpose2d(h, sd[p + ".hallucinate.0.weight"], stride=1, padding=0)
    rec("deconv1", h)
    h = swish(_bn_train(h, sd, p + ".hallucinate.1", track))
    rec("act_d1", h)
    h = F.conv_transpose2d(h, sd[p + ".hallucinate.3.weight"], stride=2, padding=1)
    rec("deconv2", h)
    h = swish(_bn_train(h, sd, p + ".hallucinate.4", track))
    rec("act_d2", h)
    h = F.conv_transpose2d(h, sd[p + ".hallucinate.6.weight"], stride=2, padding=1)
    rec("deconv3", h)
    h = swish(_bn_train(h, sd, p + ".hallucinate.7", track))
    rec("act_d3", h)
    h = F.conv_transpose2d(h, sd[p + ".hallucinate.9.weight"], stride=2, padding=1)
    rec("logits", h)
    return h


def decoder_mlp(sd, p, z):
    """pose Decoder, layer_sizes [256,512,512,7] (vae.py:121-123, 282-283, 299)."""
    h = F.relu(F.linear(z, sd[p + ".deconv_net.0.weight"], sd[p + ".deconv_net.0.bias"]))
    h = F.relu(F.linear(h, sd[p + ".deconv_net.2.weight"], sd[p + ".deconv_net.2.bias"]))
    return F.linear(h, sd[p + ".deconv_net.4.weight"], sd[p + ".deconv_net.4.bias"])


def product_of_experts(mu, logvar, eps=POE_EPS):
    """ProductOfExperts.forward (vae.py:311-318); mu/logvar: (M, B, D)."""
    var = torch.exp(logvar) + eps
    T = 1.0 / (var + eps)
    pd_mu = torch.sum(mu * T, dim=0) / torch.sum(T, dim=0)
    pd_var = 1.0 / torch.sum(T, dim=0)
    pd_logvar = torch.log(pd_var + eps)
    return pd_mu, pd_logvar


def reparametrize(mu, logvar, eps_noise):
    """Autoencoder.reparametrize (vae.py:52-61) with the N(0,1) draw passed in."""
    return eps_noise * torch.exp(0.5 * logvar) + mu


def draw_pass_noise(B, has_visual, has_tactile, latent=256, generator=None, dtype=torch.float32):
    """The reference's CPU RNG consumption for ONE MVAE.forward in training mode: the visual
    encoder's Dropout mask (vae.py:213 via :142), the tactile one (:147), then eps (:58).  On CPU
    F.dropout(p) equals empty.bernoulli_(1-p)/(1-p) under the same generator state."""
    mv = mt = None
    if has_visual:
        mv = torch.empty(B, 512, dtype=dtype).bernoulli_(1 - DROPOUT_P, generator=generator) / (1 - DROPOUT_P)
    if has_tactile:
        mt = torch.empty(B, 512, dtype=dtype).bernoulli_(1 - DROPOUT_P, generator=generator) / (1 - DROPOUT_P)
    eps = torch.randn(B, latent, generator=generator, dtype=dtype)
    return mv, mt, eps


def vae_forward(sd, x, noise, track=True, acts=None, c=None):
    """VAE.forward (vae.py:81-88): noise = (dropout_mask, eps); c = condition (CVAE) or None."""
    mask, eps = noise
    mu, lv = encoder_cnn(sd, "encoder", x, mask, track, acts, c)
    z = reparametrize(mu, lv, eps)
    if acts is not None:
        acts["z"] = z
    return decoder_cnn(sd, "decoder", z, track, acts, c), mu, lv


def mvae_forward(sd, visual, tactile, pose, noise, use_pose, track=True, acts=None, c=None):
    """MVAE.forward (vae.py:126-165).  Expert order prior, visual, tactile, pose; both image
    decoders always run (:160-161); the pose decoder runs iff use_pose (:163).  c: condition of a
    conditional=True model — used by the image encoders / decoders; the pose expert is built
    un-conditional (vae.py:118-123) and ignores it."""
    mv, mt, eps = noise
    B = (visual if visual is not None else tactile if tactile is not None else pose).size(0)
    ref = next(iter(sd.values()))
    mus = [torch.zeros(B, eps.size(1), dtype=ref.dtype)]   # prior_expert, vae.py:321-328
    lvs = [torch.zeros(B, eps.size(1), dtype=ref.dtype)]
    if visual is not None:
        m, l = encoder_cnn(sd, "visual_encoder", visual, mv, track, acts, c)
        mus.append(m), lvs.append(l)
    if tactile is not None:
        m, l = encoder_cnn(sd, "tactile_encoder", tactile, mt, track, acts, c)
        mus.append(m), lvs.append(l)
    if pose is not None and use_pose:
        m, l = encoder_mlp(sd, "pose_encoder", pose)
        mus.append(m), lvs.append(l)
    mu, lv = product_of_experts(torch.stack(mus), torch.stack(lvs))
    z = reparametrize(mu, lv, eps)
    if acts is not None:
        acts["z"] = z
    v_rec = decoder_cnn(sd, "visual_decoder", z, track, acts, c)
    t_rec = decoder_cnn(sd, "tactile_decoder", z, track, acts, c)
    p_rec = decoder_mlp(sd, "pose_decoder", z) if use_pose else None
    return v_rec, t_rec, p_rec, mu, lv


def kl_divergence(mu, lv):
    """problems.py:406, 429."""
    return -0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp())


def elbo_loss(recon_x, x, mu, lv, kl_weight, loss_mask=None):
    """Reconstruction._elbo_loss, reduce=None path (problems.py:401-419)."""
    kld = kl_divergence(mu, lv)
    if loss_mask is not None:
        bce = F.binary_cross_entropy_with_logits(recon_x.view(x.size()) * loss_mask, x * loss_mask, reduction="sum")
    else:
        bce = F.binary_cross_entropy_with_logits(recon_x.view(x.size()), x, reduction="sum")
    return (bce + kl_weight * kld) / x.size(0)


def mvae_elbo_loss(recons, targets, mu, lv, kl_weight, pose_multiplier, loss_mask=None):
    """Reconstruction._mvae_elbo_loss, reduce=None path (problems.py:421-458): images -> BCE with
    logits (sum), vectors -> MSE (sum) x pose_multiplier."""
    B = targets[0].size(0)
    err = 0
    for r, t in zip(recons, targets):
        if r.dim() > 2:
            r = r.view(t.size())
            if loss_mask is not None:
                e = F.binary_cross_entropy_with_logits(r * loss_mask, t * loss_mask, reduction="sum")
            else:
                e = F.binary_cross_entropy_with_logits(r, t, reduction="sum")
        else:
            e = pose_multiplier * F.mse_loss(r, t, reduction="sum")
        err = err + e
    return (err + kl_weight * kl_divergence(mu, lv)) / B


def elbo_per_sample(recon_x, x, mu, lv, kl_weight, loss_mask=None):
    """Reconstruction._elbo_loss with reduce=False (problems.py:408-417): element-wise BCE summed per
    sample; the KL term is the batch TOTAL added to every sample (reference behaviour)."""
    r = recon_x.view(x.size())
    if loss_mask is not None:
        bce = F.binary_cross_entropy_with_logits(r * loss_mask, x * loss_mask, reduction="none")
    else:
        bce = F.binary_cross_entropy_with_logits(r, x, reduction="none")
    return bce.sum((1, 2, 3)) + kl_weight * kl_divergence(mu, lv)


def mvae_elbo_per_sample(recons, targets, mu, lv, kl_weight, pose_multiplier, loss_mask=None):
    """Reconstruction._mvae_elbo_loss with reduce=False (problems.py:445-456): element-wise losses summed
    per sample; the KL term stays the batch TOTAL and is added to every sample (reference behaviour)."""
    err = 0
    for r, t in zip(recons, targets):
        if r.dim() > 2:
            r = r.view(t.size())
            if loss_mask is not None:
                e = F.binary_cross_entropy_with_logits(r * loss_mask, t * loss_mask, reduction="none").sum((1, 2, 3))
            else:
                e = F.binary_cross_entropy_with_logits(r, t, reduction="none").sum((1, 2, 3))
        else:
            e = pose_multiplier * F.mse_loss(r, t, reduction="none").sum(1)
        err = err + e
    return err + kl_weight * kl_divergence(mu, lv)


MVAE_PASSES_NOPOSE = [(True, True, False), (True, False, False), (False, True, False)]
MVAE_PASSES_POSE = MVAE_PASSES_NOPOSE + [(True, True, True), (True, False, True), (False, True, True),
                                         (False, False, True)]


def evaluate_mvae(sd, x, targets, kl_weight, pose_multiplier, use_pose, noises, loss_mask=None, track=True,
                  acts=None, condition=None):
    """Reconstruction._evaluate_mvae (problems.py:473-546): sub-sampled training objective, the sum
    of 3 (or 7 with pose) ELBOs.  x = [visual, tactile(, pose)], targets likewise; noises = one
    (mask_v, mask_t, eps) triple per pass in pass order.  Returns (outputs, loss, per_pass) with the
    reference's quirky `outputs` bindings (recon_x of the joint pass, means/log_var of the LAST pass)."""
    passes = MVAE_PASSES_POSE if use_pose else MVAE_PASSES_NOPOSE
    loss = 0
    per_pass = []
    rec_joint = None
    perf = {}
    for i, ((hv, ht, hp), noise) in enumerate(zip(passes, noises)):
        a = {} if acts is not None else None
        v_rec, t_rec, p_rec, mu, lv = mvae_forward(sd, x[0] if hv else None, x[1] if ht else None,
                                                   x[2] if hp else None, noise, use_pose, track, a, condition)
        if acts is not None:
            acts[i] = a
        recs, tgts = [], []
        if hv:
            recs.append(v_rec), tgts.append(targets[0])
        if ht:
            recs.append(t_rec), tgts.append(targets[1])
        if hp:
            recs.append(p_rec), tgts.append(targets[2])
        l = mvae_elbo_loss(recs, tgts, mu, lv, kl_weight, pose_multiplier, loss_mask)
        loss = loss + l
        per_pass.append({"loss": l, "mu": mu, "lv": lv, "v_rec": v_rec, "t_rec": t_rec, "p_rec": p_rec})
        with torch.no_grad():
            if (hv, ht, hp) == (True, False, False):   # problems.py:499-501
                perf["visual"] = F.binary_cross_entropy_with_logits(v_rec, targets[0], reduction="mean").item()
            if (hv, ht, hp) == (False, True, False):   # problems.py:502-503
                perf["tactile"] = F.binary_cross_entropy_with_logits(t_rec, targets[1], reduction="mean").item()
            if (hv, ht, hp) == (False, False, True):   # problems.py:535
                perf["pose"] = F.mse_loss(p_rec, targets[2], reduction="mean").item()
        if (hv, ht, hp) == ((True, True, True) if use_pose else (True, True, False)):
            rec_joint = [v_rec, t_rec] + ([p_rec] if use_pose else [])
    outputs = {"recon_x": rec_joint, "means": per_pass[-1]["mu"], "log_var": per_pass[-1]["lv"],
               "perf_measure": perf}
    return outputs, loss, per_pass


def regressor_forward(sd, x, dropout_mask=None, c=None, track=True, acts=None):
    """Regressor.forward (models.py:64-77): conv_net -> flatten -> fc_net (Linear, Swish, Dropout) ->
    [cat condition] -> out_net (Linear ReLU Linear ReLU Linear).  The conv_net / fc_net stack is the cnn
    Encoder's (models.py:38-55 == vae.py:198-214), restated once more here because the parameter names
    carry no sub-module prefix."""
    def rec(k, v):
        if acts is not None:
            acts[k] = v
    h = swish(F.conv2d(x, sd["conv_net.0.weight"], stride=2, padding=1))
    h = F.conv2d(h, sd["conv_net.2.weight"], stride=2, padding=1)
    h = swish(_bn_train(h, sd, "conv_net.3", track))
    h = F.conv2d(h, sd["conv_net.5.weight"], stride=2, padding=1)
    h = swish(_bn_train(h, sd, "conv_net.6", track))
    h = F.conv2d(h, sd["conv_net.8.weight"], stride=1, padding=0)
    h = swish(_bn_train(h, sd, "conv_net.9", track))
    rec("act4", h)
    h = swish(F.linear(h.reshape(h.size(0), -1), sd["fc_net.0.weight"], sd["fc_net.0.bias"]))
    if dropout_mask is not None:
        h = h * dropout_mask
    rec("fc", h)
    h = _cat_condition(h, c)
    h = F.relu(F.linear(h, sd["out_net.0.weight"], sd["out_net.0.bias"]))
    rec("a1", h)
    h = F.relu(F.linear(h, sd["out_net.2.weight"], sd["out_net.2.bias"]))
    return F.linear(h, sd["out_net.4.weight"], sd["out_net.4.bias"])


def evaluate_regression(sd, x, target, dropout_mask, condition=None, track=True, acts=None):
    """Regression._evaluate_model (problems.py:321-332): MSE 'sum' loss, mean-MSE metric."""
    out = regressor_forward(sd, x, dropout_mask, condition, track, acts)
    loss = F.mse_loss(out.view(target.size()), target, reduction="sum")
    with torch.no_grad():
        m = F.mse_loss(out.view(target.size()), target, reduction="mean").item()
    return {"outputs": out, "perf_measure": {"pose": m}}, loss


def regression_parse_input(data, target, seq_length, input_type):
    """Regression.parse_input (problems.py:291-316) on seq-collated lists."""
    L = seq_length
    k = 0 if input_type == "visual" else 1
    shock = data[4][::L] if len(data) > 4 else None
    return {"model_input": data[k][::L], "shock": shock}, target[2][::L]


def evaluate_vae(sd, x, target, kl_weight, noise, loss_mask=None, track=True, acts=None, input_type="visual",
                 condition=None):
    """SeqModeling._evaluate_model, plain VAE / CVAE branch (problems.py:702-716)."""
    recon, mu, lv = vae_forward(sd, x, noise, track, acts, condition)
    loss = elbo_loss(recon, target, mu, lv, kl_weight, loss_mask)
    with torch.no_grad():
        m = F.binary_cross_entropy_with_logits(recon.view(target.size()), target, reduction="mean").item()
    return {"recon_x": recon, "means": mu, "log_var": lv, "perf_measure": {input_type: m}}, loss


def anneal_kl(epoch, annealing_epochs):
    """Problem._anneal_KL (problems.py:212-216)."""
    return (epoch + 1) / annealing_epochs if epoch < annealing_epochs else 1


def seq_parse_input(data, target, seq_length, input_type):
    """SeqModeling.parse_input (problems.py:634-673) without the device moves: first frame of each
    sequence via [::L].  Returns (inputs dict, targets dict)."""
    L = seq_length
    if input_type == "visual":
        mi, to = data[0][::L], target[0][::L]
    elif input_type == "tactile":
        mi, to = data[1][::L], target[1][::L]
    else:
        mi, to = [data[0][::L], data[1][::L]], [target[0][::L], target[1][::L]]
    inputs = {"model_input": mi, "input_object_pose": None, "input_available_modals": None, "shock": None}
    targets = {"target_output": to, "target_object_pose": None, "loss_mask": None}
    if len(data) > 2:
        inputs["input_object_pose"] = [data[2][::L]]
        inputs["input_available_modals"] = data[3][::L]
        targets["target_object_pose"] = [target[2][::L]]
        targets["loss_mask"] = target[3][::L]
        inputs["shock"] = data[4][::L] if len(data) > 4 else None
    return inputs, targets


def dyn_parse_input(data, target, seq_length, input_type):
    """DynModeling.parse_input (problems.py:765-803): one-step dynamics targets = roll(-1) with every
    last-of-sequence row replaced by the resting-state target; the POSE target is a bare roll with
    no fix-up (:798), so the last pose of a sequence targets the next sequence's first pose and the
    very last row wraps to row 0 — reproduced on purpose."""
    L = seq_length

    def shifted(i):
        t = torch.roll(data[i], -1, dims=0).clone()
        t[L - 1::L] = target[i][L - 1::L]
        return t
    if input_type == "visual":
        mi, to = data[0], shifted(0)
    elif input_type == "tactile":
        mi, to = data[1], shifted(1)
    else:
        mi, to = [data[0], data[1]], [shifted(0), shifted(1)]
    inputs = {"model_input": mi, "input_object_pose": [data[2]], "input_available_modals": data[3],
              "shock": data[4] if len(data) > 4 else None}
    targets = {"target_output": to, "target_object_pose": [torch.roll(data[2], -1, dims=0)],
               "loss_mask": target[3]}
    return inputs, targets


def adam_step(params, grads, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam defaults as used at problems.py:138, one step, in place.
    state: dict with 'step', 'm' (list), 'v' (list)."""
    state["step"] += 1
    t = state["step"]
    b1, b2 = betas
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    for p, g, m, v in zip(params, grads, state["m"], state["v"]):
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


def train_step(sd, param_keys, problem, batch, kl_weight, pose_multiplier, noises, adam_state, lr=1e-3,
               loss_mask=None, condition=None):
    """One iteration of Problem._train_epoch's body (problems.py:150-155): zero_grad, evaluate,
    backward, Adam.  `problem` is 'vae' or 'mvae' / 'mvae+pose'.  Returns (outputs, loss, grads)."""
    params = [sd[k].requires_grad_(True) for k in param_keys]
    for p in params:
        p.grad = None
    if problem == "vae":
        outputs, loss = evaluate_vae(sd, batch["x"], batch["target"], kl_weight, noises[0], loss_mask,
                                     condition=condition)
    else:
        outputs, loss, _ = evaluate_mvae(sd, batch["x"], batch["targets"], kl_weight, pose_multiplier,
                                         problem == "mvae+pose", noises, loss_mask, condition=condition)
    loss.backward()
    grads = [p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p) for p in params]
    with torch.no_grad():
        for p in params:
            p.requires_grad_(False)
        adam_step(params, grads, adam_state, lr=lr)
    return outputs, loss.detach(), grads
