"""End-to-end parity of the B200 path against the oracle (CPU fp32 restatement of the reference,
pinned to the live reference by tests/golden/) on identical weights, inputs and noise.

Stated tolerances (north_star): conv-path tensors (fp16 operands = TF32-equivalent mantissa,
fp32 accumulate) rel 1e-3 per layer; fp32 PoE/KL 1e-5 given identical inputs (kernel test);
end-to-end quantities are compared with the norm-wise relative error  ||a-b|| / ||b||  and the
thresholds written next to each assert.  Masks / index work is bit-exact.
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mmdyn_oracle as orc  # noqa: E402

DEV = "cuda"
KW = dict(condition_dim=0, input_dim=4096, architecture="cnn", conditional=False, categorical_conditions=False,
          latent_size=256)


def nrel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nchw(t):
    return t.permute(0, 3, 1, 2)


def make(model_name, use_pose=False, seed=0):
    from mmdyn_b200.pytorch.models.models import setup_model
    torch.manual_seed(seed)
    kw = dict(KW)
    if "mvae" in model_name:
        kw["use_pose"] = use_pose
    m = setup_model(model_name, cross_modal="mvae" in model_name, **kw)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    return m.to(DEV), sd


def batch(B, seed=1):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    return dict(v=r(B, 3, 64, 64), t=r(B, 3, 64, 64), p=r(B, 7), tv=r(B, 3, 64, 64), tt=r(B, 3, 64, 64), tp=r(B, 7))


def oracle_noises(passes, B, seed):
    g = torch.Generator().manual_seed(seed)
    return [orc.draw_pass_noise(B, hv, ht, generator=g) for (hv, ht, hp) in passes]


@pytest.mark.parametrize("use_pose", [False, True])
def test_mvae_fused_step_matches_oracle(use_pose):
    from mmdyn_b200 import engine, noise, optim
    B = 8
    model, sd = make("cnn-mvae", use_pose)
    d = batch(B)
    passes = orc.MVAE_PASSES_POSE if use_pose else orc.MVAE_PASSES_NOPOSE
    klw, pm = 0.02, 1000.0
    # ---- oracle: forward, backward, one Adam step ----
    pkeys = [k for k, _ in model.named_parameters()]
    x_o = [d["v"], d["t"]] + ([d["p"]] if use_pose else [])
    t_o = [d["tv"], d["tt"]] + ([d["tp"]] if use_pose else [])
    acts = {}
    sd_o = copy.deepcopy(sd)
    for k in pkeys:
        sd_o[k].requires_grad_(True)
    out_o, loss_o, per_pass = orc.evaluate_mvae(sd_o, x_o, t_o, klw, pm, use_pose, oracle_noises(passes, B, 7),
                                                acts=acts)
    loss_o.backward()
    grads_o = {k: sd_o[k].grad.clone() for k in pkeys}
    # ---- B200 path ----
    eng = engine.StepEngine(model, "mvae", use_pose=use_pose, pose_multiplier=pm,
                            noise_src=noise.HostNoise(torch.Generator().manual_seed(7)), exact_running_stats=True)
    opt = optim.FusedAdam(model, lr=1e-3)
    opt.zero_grad()
    x_d = [t.to(DEV) for t in x_o]
    t_d = [t.to(DEV) for t in t_o]
    out_d, loss_d = eng.evaluate(x_d, t_d, klw)
    torch.cuda.synchronize()
    # per-layer activations of the shared visual trunk and of the visual decoder (joint pass)
    rec = eng._state["enc_rec"]["v"]
    a0 = acts[0]
    errs = {
        "enc.conv1": nrel(nchw(rec["raw1"]), a0["visual_encoder.conv1"]),
        "enc.conv2": nrel(nchw(rec["raw2"]), a0["visual_encoder.conv2"]),
        "enc.act2": nrel(nchw(rec["act2"]), a0["visual_encoder.act2"]),
        "enc.conv3": nrel(nchw(rec["raw3"]), a0["visual_encoder.conv3"]),
        "enc.conv4": nrel(nchw(rec["raw4"]), a0["visual_encoder.conv4"]),
        "enc.act4": nrel(nchw(rec["act4"]), a0["visual_encoder.act4"]),
    }
    dr = eng._state["dec_rec"]["v"]
    errs.update({
        "dec.up": nrel(nchw(dr["act0"][:B]), a0["visual_decoder.up"]),
        "dec.deconv1": nrel(nchw(dr["raw1"][:B]), a0["visual_decoder.deconv1"]),
        "dec.deconv2": nrel(nchw(dr["raw2"][:B]), a0["visual_decoder.deconv2"]),
        "dec.deconv3": nrel(nchw(dr["raw3"][:B]), a0["visual_decoder.deconv3"]),
    })
    # logits are only materialised for the pass `outputs` returns (the joint pass); the others live
    # only inside the fused loss epilogue.  exact_running_stats: decoder group index == pass index.
    jg = 3 if use_pose else 0
    errs["dec.logits"] = nrel(dr["logits"][jg * B:(jg + 1) * B], acts[jg]["visual_decoder.logits"])
    mu_d, lv_d = eng.ws.bufs["mu"], eng.ws.bufs["lv"]
    for i, pp in enumerate(per_pass):
        errs[f"mu[{i}]"] = nrel(mu_d[i], pp["mu"])
        errs[f"lv[{i}]"] = nrel(lv_d[i], pp["lv"])
    errs["loss"] = abs(loss_d.item() - loss_o.item()) / abs(loss_o.item())
    print("\n".join(f"{k:14s} {v:.3e}" for k, v in errs.items()))
    for k, v in errs.items():
        # chained activations compound the per-layer 1e-3 budget over depth (isolated layers are
        # held to 1e-3 in test_layers_in_isolation); the scalar loss averages the noise out
        # measured (r2 log, profiles/r2_parity_measured.txt): activations <= 1.28e-3, mu <= 1.37e-3, loss 6e-7
        tol = 1e-5 if k == "loss" else 2e-3
        assert v < tol, (k, v, tol)
    # outputs dict: reference bindings (recon_x of the joint pass, posterior of the last pass)
    for a, b in zip(out_d["recon_x"], out_o["recon_x"]):
        assert nrel(a, b) < 2e-3
    assert nrel(out_d["means"], out_o["means"]) < 2e-3 and nrel(out_d["log_var"], out_o["log_var"]) < 2e-3
    for k, v in out_o["perf_measure"].items():
        assert abs(float(out_d["perf_measure"][k]) - v) / abs(v) < 1e-3, k
    # BatchNorm running statistics and counters (exact_running_stats=True replays every pass)
    sdd = model.state_dict()
    for k in sd_o:
        if k.endswith("num_batches_tracked"):
            assert int(sdd[k]) == int(sd_o[k]), k
        elif "running_" in k:
            assert nrel(sdd[k], sd_o[k]) < 2e-3, (k, nrel(sdd[k], sd_o[k]))
    # ---- backward ----
    loss_d.backward()
    torch.cuda.synchronize()
    gerr = {k: nrel(p.grad, grads_o[k]) for k, p in model.named_parameters()}
    worst = sorted(gerr.items(), key=lambda kv: -kv[1])[:8]
    print("worst gradient errors:\n" + "\n".join(f"{k:50s} {v:.3e}" for k, v in worst))
    # fp16 activations/gradients through 10 layers: norm-wise 1e-2 per tensor, 3e-3 on the whole arena.
    # With the pose expert the ReLU MLP decoder sits behind z: a 1e-3 perturbation of z (fp16 image
    # trunks) flips a few ReLU units of the 8-row batch and moves dz of that pass by ~2 % (measured:
    # the pose kernels themselves reproduce torch to 4e-7 on identical inputs), so the bound is 5e-2.
    # measured (r2): worst tensor 2.6e-3 without / 1.34e-2 with the pose expert; the pose figure is the ReLU-flip
    # effect isolated row by row in tests/test_parity2_gpu.py::test_pose_gradient_error_is_explained_by_relu_flips
    tol_t, tol_all = (4e-2, 2e-2) if use_pose else (6e-3, 3e-3)
    for k, v in gerr.items():
        assert v < tol_t, (k, v)
    flat_o = torch.cat([grads_o[k].reshape(-1) for k in pkeys])
    flat_d = torch.cat([p.grad.reshape(-1) for _, p in model.named_parameters()])
    assert nrel(flat_d, flat_o) < tol_all, nrel(flat_d, flat_o)
    # ---- one Adam step (problems.py:155) ----
    # (a) in situ: the fused update applied to the B200 gradients equals the oracle's Adam applied to
    #     the SAME gradients to fp32 round-off (1e-6 abs on parameters of scale 1e-2..1);
    # (b) against the oracle's own step: step 1 of Adam is lr*g/(|g|+eps) ~ lr*sign(g), so entries
    #     whose gradient is below the fp16 noise floor may flip sign; bound the relative L2 distance of
    #     the parameter DELTAS (0.15) instead of an element-wise tolerance.
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    g_dev = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters()}
    opt.step()
    torch.cuda.synchronize()
    with torch.no_grad():
        def oracle_adam(grads):
            st = {"step": 0, "m": [torch.zeros_like(sd[k]) for k in pkeys], "v": [torch.zeros_like(sd[k]) for k in pkeys]}
            params = [sd[k].clone() for k in pkeys]
            orc.adam_step(params, [grads[k] for k in pkeys], st, lr=1e-3)
            return params
        same_g = oracle_adam(g_dev)
        own_g = oracle_adam(grads_o)
    num = den = 0.0
    for (k, p), ps, po in zip(model.named_parameters(), same_g, own_g):
        assert (p.detach().cpu() - ps).abs().max().item() < 1e-6, k
        dd, do = (p.detach().cpu() - before[k].cpu()).double(), (po - sd[k]).double()
        num += (dd - do).pow(2).sum().item()
        den += do.pow(2).sum().item()
    assert (num / den) ** 0.5 < 0.15, (num / den) ** 0.5


def test_vae_fused_step_and_module_api_match_oracle():
    from mmdyn_b200 import engine, noise
    B = 8
    model, sd = make("cnn-vae")
    d = batch(B, seed=3)
    klw = 0.02
    pkeys = [k for k, _ in model.named_parameters()]
    g = torch.Generator().manual_seed(11)
    mask, _, eps = orc.draw_pass_noise(B, True, False, generator=g)
    sd_o = copy.deepcopy(sd)
    for k in pkeys:
        sd_o[k].requires_grad_(True)
    out_o, loss_o = orc.evaluate_vae(sd_o, d["t"], d["tt"], klw, (mask, eps), track=True, input_type="tactile")
    loss_o.backward()
    # fused step
    eng = engine.StepEngine(model, "vae", noise_src=noise.HostNoise(torch.Generator().manual_seed(11)))
    out_d, loss_d = eng.evaluate(d["t"].to(DEV), d["tt"].to(DEV), klw)
    loss_d.backward()
    torch.cuda.synchronize()
    assert abs(loss_d.item() - loss_o.item()) / loss_o.item() < 1e-4
    assert nrel(out_d["recon_x"], out_o["recon_x"]) < 2e-3
    assert nrel(out_d["means"], out_o["means"]) < 2e-3 and nrel(out_d["log_var"], out_o["log_var"]) < 2e-3
    assert abs(float(out_d["perf_measure"]["x"]) - out_o["perf_measure"]["tactile"]) / out_o["perf_measure"]["tactile"] < 1e-3
    g1 = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    for k in pkeys:
        assert nrel(g1[k], sd_o[k].grad) < 1e-2, (k, nrel(g1[k], sd_o[k].grad))
    # module-level API (model(x) + torch loss + autograd) on a fresh copy of the same weights
    model2, _ = make("cnn-vae")
    model2.noise = noise.HostNoise(torch.Generator().manual_seed(11))
    recon, mu, lv = model2(d["t"].to(DEV))
    kld = -0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp())
    bce = torch.nn.functional.binary_cross_entropy_with_logits(recon, d["tt"].to(DEV), reduction="sum")
    loss2 = (bce + klw * kld) / B
    loss2.backward()
    torch.cuda.synchronize()
    assert abs(loss2.item() - loss_o.item()) / loss_o.item() < 1e-4
    assert nrel(mu, out_o["means"]) < 2e-3
    for k, p in model2.named_parameters():
        assert nrel(p.grad, sd_o[k].grad) < 1e-2, (k, nrel(p.grad, sd_o[k].grad))


def test_mvae_module_api_forward_and_inference():
    from mmdyn_b200 import noise
    B = 4
    model, sd = make("cnn-mvae", True, seed=2)
    model.noise = noise.HostNoise(torch.Generator().manual_seed(5))
    d = batch(B, seed=9)
    g = torch.Generator().manual_seed(5)
    nz = orc.draw_pass_noise(B, True, False, generator=g)
    sd_o = copy.deepcopy(sd)
    v_o, t_o, p_o, mu_o, lv_o = orc.mvae_forward(sd_o, d["v"], None, d["p"], nz, True)
    v, t, p, mu, lv = model([d["v"].to(DEV), None], pose=d["p"].to(DEV))
    torch.cuda.synchronize()
    assert nrel(mu, mu_o) < 2e-3 and nrel(lv, lv_o) < 2e-3
    assert nrel(v, v_o) < 2e-3 and nrel(t, t_o) < 2e-3 and nrel(p, p_o) < 2e-3
    vi, ti = model.inference(n=6)
    assert vi.shape == (6, 3, 64, 64) and ti.shape == (6, 3, 64, 64) and torch.isfinite(vi).all()
    # state_dict round trip keeps the reference's 106 keys and survives a re-load
    keys = list(model.state_dict().keys())
    assert len(keys) == 106 and keys[0] == "visual_encoder.conv_net.0.weight"
    model.load_state_dict({k: v_.to(DEV) for k, v_ in sd.items()})
    v2, *_ = model([d["v"].to(DEV), None], pose=d["p"].to(DEV))
    assert torch.isfinite(v2).all()


def test_product_of_experts_module_matches_reference_formula():
    from mmdyn_b200.pytorch.models.vae import ProductOfExperts
    torch.manual_seed(0)
    mu, lv = torch.randn(4, 9, 256), torch.randn(4, 9, 256) * 0.5
    mu[0].zero_(), lv[0].zero_()
    m_o, l_o = orc.product_of_experts(mu.double(), lv.double())
    m_d, l_d = ProductOfExperts()(mu.to(DEV), lv.to(DEV))
    assert nrel(m_d, m_o) < 1e-5 and nrel(l_d, l_o) < 1e-5


def test_layers_in_isolation():
    """Per-layer parity at the stated 1e-3 (norm-wise): every tensor-core layer of the model is fed the
    ORACLE's fp32 input activation of that layer (so errors do not compound) with the model's real
    weights, and compared with the oracle's output of the same layer."""
    from mmdyn_b200 import engine, ops
    B = 6
    model, sd = make("cnn-vae", seed=4)
    d = batch(B, seed=6)
    g = torch.Generator().manual_seed(3)
    _, _, eps = orc.draw_pass_noise(B, False, False, generator=g)
    acts = {}
    orc.vae_forward(copy.deepcopy(sd), d["v"], (None, eps), track=False, acts=acts)
    arena, ex = engine.get_execs(model, torch.device(DEV))
    enc, dec = ex["enc"]["encoder"], ex["dec"]["decoder"]
    enc.refresh()

    def nhwc16(t):
        return t.permute(0, 2, 3, 1).contiguous().half().to(DEV)

    def run(pl, a_in, out_shape, bias=None, dtype=torch.float16):
        out = torch.zeros(out_shape, dtype=dtype, device=DEV)
        ops.igemm(pl.lp.fwd, a_in, pl.Wf, out, B, bias=bias, out_mode=1 if dtype == torch.float32 and pl.lp.fwd.out_mode == 0 else None)
        torch.cuda.synchronize()
        return out

    errs = {}
    e = "encoder."
    raw1 = torch.zeros(B, 32, 32, 32, dtype=torch.float16, device=DEV)
    ops.conv1_fwd(d["v"].to(DEV), enc.c1.Wf, raw1, B)
    errs["conv1"] = nrel(nchw(raw1), acts[e + "conv1"])
    errs["conv2"] = nrel(nchw(run(enc.c2, nhwc16(orc.swish(acts[e + "conv1"])), (B, 16, 16, 64))), acts[e + "conv2"])
    errs["conv3"] = nrel(nchw(run(enc.c3, nhwc16(acts[e + "act2"]), (B, 8, 8, 128))), acts[e + "conv3"])
    errs["conv4"] = nrel(nchw(run(enc.c4, nhwc16(acts[e + "act3"]), (B, 5, 5, 256))), acts[e + "conv4"])
    fc = run(enc.fc, nhwc16(acts[e + "act4"]).reshape(B, 6400), (B, 512), bias=enc.fc.bias, dtype=torch.float32)
    errs["fc"] = nrel(orc.swish(fc.cpu()), acts[e + "fc"])
    heads = run(enc.heads, acts[e + "fc"].half().to(DEV), (B, 512), bias=enc.heads.bias, dtype=torch.float32)
    errs["heads"] = nrel(heads, torch.cat([acts[e + "mu"], acts[e + "lv"]], 1))
    dd = "decoder."
    up = run(dec.up, acts["z"].half().to(DEV), (B, 5, 5, 256), bias=dec.up.bias)
    errs["upsample"] = nrel(orc.swish(nchw(up).float().cpu()), acts[dd + "up"])
    errs["deconv1"] = nrel(nchw(run(dec.d1, nhwc16(acts[dd + "up"]), (B, 8, 8, 128))), acts[dd + "deconv1"])
    errs["deconv2"] = nrel(nchw(run(dec.d2, nhwc16(acts[dd + "act_d1"]), (B, 16, 16, 64))), acts[dd + "deconv2"])
    errs["deconv3"] = nrel(nchw(run(dec.d3, nhwc16(acts[dd + "act_d2"]), (B, 32, 32, 32))), acts[dd + "deconv3"])
    errs["deconv4"] = nrel(run(dec.d4, nhwc16(acts[dd + "act_d3"]), (B, 3, 64, 64), dtype=torch.float32), acts[dd + "logits"])
    print("\n".join(f"isolated {k:10s} {v:.3e}" for k, v in errs.items()))
    for k, v in errs.items():
        assert v < 1e-3, (k, v)


def _dataset_free_problem(model, use_pose, klw=0.02):
    from mmdyn_b200.pytorch.problems import problems
    pr = object.__new__(problems.SeqModeling)
    pr.parameters = {"model_name": "cnn-mvae", "input_type": "visuotactile", "use_pose": use_pose, "mask_loss": False}
    pr._kl_weight, pr._pose_multiplier, pr._conditional = klw, 1000.0, False
    pr._model, pr._cross_modal, pr._engine = model, True, None
    return pr


def test_elbo_entry_points_and_pass_loop_match_oracle_and_fused_step():
    """Reconstruction._mvae_elbo_loss / _elbo_loss called on their own (problems.py:401-458), the
    pass-by-pass evaluator built on them, the per-sample scoring path (reduce=False) and the fused
    step all agree with the oracle on the same weights / inputs / noise."""
    from mmdyn_b200 import noise
    B, klw = 4, 0.02
    model, sd = make("cnn-mvae", True, seed=5)
    d = batch(B, seed=8)
    x_o, t_o = [d["v"], d["t"], d["p"]], [d["tv"], d["tt"], d["tp"]]
    x_d, t_d = [a.to(DEV) for a in x_o], [a.to(DEV) for a in t_o]
    out_o, loss_o, per_pass = orc.evaluate_mvae(copy.deepcopy(sd), x_o, t_o, klw, 1000.0, True,
                                                oracle_noises(orc.MVAE_PASSES_POSE, B, 21))
    pr = _dataset_free_problem(model, True, klw)
    # (a) pass loop through the module-level API + stand-alone loss kernels, with autograd
    model.noise = noise.HostNoise(torch.Generator().manual_seed(21))
    out_p, loss_p = pr._evaluate_mvae_passes(x_d, t_d)
    assert abs(loss_p.item() - loss_o.item()) / loss_o.item() < 1e-4
    loss_p.backward()
    g_pass = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    # (b) fused step on the same noise
    for p in model.parameters():
        p.grad = None
    model.noise = noise.HostNoise(torch.Generator().manual_seed(21))
    out_f, loss_f = pr._evaluate_model({"model_input": x_d[:2], "input_object_pose": [x_d[2]]},
                                       {"target_output": t_d[:2], "target_object_pose": [t_d[2]], "loss_mask": None})
    assert abs(loss_f.item() - loss_p.item()) / loss_p.item() < 1e-5
    loss_f.backward()
    torch.cuda.synchronize()
    num = sum((p.grad - g_pass[k]).double().pow(2).sum().item() for k, p in model.named_parameters())
    den = sum(g_pass[k].double().pow(2).sum().item() for k in g_pass)
    assert (num / den) ** 0.5 < 2e-2, (num / den) ** 0.5  # two independent backward implementations
    # (c) per-sample scoring, reduce=False
    with torch.no_grad():
        model.noise = noise.HostNoise(torch.Generator().manual_seed(21))
        _, score = pr._evaluate_mvae_passes(x_d, t_d, reduce=False)
    ref = 0
    for (hv, ht, hp), pp in zip(orc.MVAE_PASSES_POSE, per_pass):
        recs = ([pp["v_rec"]] if hv else []) + ([pp["t_rec"]] if ht else []) + ([pp["p_rec"]] if hp else [])
        tgts = ([t_o[0]] if hv else []) + ([t_o[1]] if ht else []) + ([t_o[2]] if hp else [])
        ref = ref + orc.mvae_elbo_per_sample(recs, tgts, pp["mu"], pp["lv"], klw, 1000.0)
    assert score.shape == (B,) and nrel(score, ref) < 1e-4
    # (d) VAE entry point with a loss mask
    recon, mu, lv = torch.randn(B, 3, 64, 64, device=DEV), torch.randn(B, 256, device=DEV), torch.randn(B, 256, device=DEV) * 0.3
    mask = (torch.rand(B, 3, 64, 64, device=DEV) > 0.5).float()
    got = pr._elbo_loss(recon, t_d[0], mu, lv, loss_mask=mask)
    want = orc.elbo_loss(recon.cpu(), t_o[0], mu.cpu(), lv.cpu(), klw, mask.cpu())
    assert abs(got.item() - want.item()) / abs(want.item()) < 1e-5


def test_data_parallel_semantics_on_one_gpu():
    """Parity definition of the data-parallel path: N ranks == the oracle run on each
    shard with identical weights, gradients averaged.  Emulated on one GPU: two shards through the fused
    step, gradients summed in the arena, `grad_prescale = 1/2` in the fused Adam."""
    from mmdyn_b200 import engine, noise, optim, parallel
    B, klw = 8, 0.02
    model, sd = make("cnn-mvae", False, seed=9)
    d = batch(B, seed=2)
    pkeys = [k for k, _ in model.named_parameters()]
    shards = [parallel.shard_rows(B, 2, r) for r in range(2)]
    assert shards == [(0, 4), (4, 8)]
    grads_o = None
    for r, (a, b) in enumerate(shards):
        sd_o = copy.deepcopy(sd)
        for k in pkeys:
            sd_o[k].requires_grad_(True)
        _, loss_o, _ = orc.evaluate_mvae(sd_o, [d["v"][a:b], d["t"][a:b]], [d["tv"][a:b], d["tt"][a:b]], klw, 1000.0,
                                         False, oracle_noises(orc.MVAE_PASSES_NOPOSE, b - a, 30 + r))
        loss_o.backward()
        g = [sd_o[k].grad.clone() for k in pkeys]
        grads_o = g if grads_o is None else [x + y for x, y in zip(grads_o, g)]
    grads_o = [g / 2 for g in grads_o]
    opt = optim.FusedAdam(model, lr=1e-3)
    opt.grad_prescale = 0.5
    opt.zero_grad()
    for r, (a, b) in enumerate(shards):
        eng = engine.StepEngine(model, "mvae", noise_src=noise.HostNoise(torch.Generator().manual_seed(30 + r)))
        _, loss = eng.evaluate([d["v"][a:b].to(DEV), d["t"][a:b].to(DEV)], [d["tv"][a:b].to(DEV), d["tt"][a:b].to(DEV)], klw)
        loss.backward()  # accumulates into the gradient arena (the all-reduce SUM of a 2-rank job)
    torch.cuda.synchronize()
    flat_d = torch.cat([0.5 * p.grad.reshape(-1) for _, p in model.named_parameters()])
    flat_o = torch.cat([g.reshape(-1) for g in grads_o])
    assert nrel(flat_d, flat_o) < 5e-3, nrel(flat_d, flat_o)
    before = torch.cat([p.detach().reshape(-1).clone() for p in model.parameters()])
    opt.step()
    after = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.isfinite(after).all() and (after - before).abs().max().item() <= 1.0001e-3  # |lr * m/sqrt(v)| <= lr


# ---------------------------------------------------------------------------------------------
# CVAE shock conditioning (vae.py:196, 231-237, 257, 286-291)
# ---------------------------------------------------------------------------------------------
def make_cond(model_name, use_pose=False, seed=0, cd=3):
    from mmdyn_b200.pytorch.models.models import setup_model
    torch.manual_seed(seed)
    kw = dict(KW, condition_dim=cd, conditional=True)
    if "mvae" in model_name:
        kw["use_pose"] = use_pose
    m = setup_model(model_name, cross_modal="mvae" in model_name, **kw)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    return m.to(DEV), sd


def test_cvae_fused_step_and_module_api_match_oracle():
    """cnn-vae --conditional: heads Linear(515,256) and upsample Linear(259,6400); the condition columns
    are a rank-3 fp32 term next to the tensor-core GEMM.  Loss, posterior, reconstruction and every
    parameter gradient (the condition columns of the three weights checked on their own)."""
    from mmdyn_b200 import engine, noise
    B, klw = 8, 0.02
    model, sd = make_cond("cnn-vae")
    assert tuple(sd["encoder.linear_means.weight"].shape) == (256, 515)
    assert tuple(sd["decoder.upsample.0.weight"].shape) == (6400, 259)
    d = batch(B, seed=3)
    c = 2 * torch.rand(B, 3, generator=torch.Generator().manual_seed(17)) - 1
    pkeys = [k for k, _ in model.named_parameters()]
    mask, _, eps = orc.draw_pass_noise(B, True, False, generator=torch.Generator().manual_seed(11))
    sd_o = copy.deepcopy(sd)
    for k in pkeys:
        sd_o[k].requires_grad_(True)
    out_o, loss_o = orc.evaluate_vae(sd_o, d["v"], d["tv"], klw, (mask, eps), track=True, input_type="visual",
                                     condition=c)
    loss_o.backward()
    # the condition must matter, otherwise this test proves nothing
    _, loss_nc = orc.evaluate_vae(copy.deepcopy(sd), d["v"], d["tv"], klw, (mask, eps), track=False,
                                  input_type="visual", condition=torch.zeros_like(c))
    assert abs(loss_nc.item() - loss_o.item()) / loss_o.item() > 1e-5
    eng = engine.StepEngine(model, "vae", noise_src=noise.HostNoise(torch.Generator().manual_seed(11)))
    with pytest.raises(ValueError, match="condition"):
        eng.evaluate(d["v"].to(DEV), d["tv"].to(DEV), klw)
    eng = engine.StepEngine(model, "vae", noise_src=noise.HostNoise(torch.Generator().manual_seed(11)))
    out_d, loss_d = eng.evaluate(d["v"].to(DEV), d["tv"].to(DEV), klw, condition=c.to(DEV))
    loss_d.backward()
    torch.cuda.synchronize()
    assert abs(loss_d.item() - loss_o.item()) / loss_o.item() < 1e-4
    assert nrel(out_d["recon_x"], out_o["recon_x"]) < 2e-3
    assert nrel(out_d["means"], out_o["means"]) < 2e-3 and nrel(out_d["log_var"], out_o["log_var"]) < 2e-3
    for k, p in model.named_parameters():
        assert nrel(p.grad, sd_o[k].grad) < 1e-2, (k, nrel(p.grad, sd_o[k].grad))
    for k, col0 in (("encoder.linear_means.weight", 512), ("encoder.linear_log_var.weight", 512),
                    ("decoder.upsample.0.weight", 256)):
        e = nrel(dict(model.named_parameters())[k].grad[:, col0:], sd_o[k].grad[:, col0:])
        print(f"condition columns of {k}: {e:.3e}")
        assert e < 1e-2, (k, e)
    # module-level API: model(x, c) + torch loss + autograd on a fresh copy of the same weights
    model2, _ = make_cond("cnn-vae")
    model2.noise = noise.HostNoise(torch.Generator().manual_seed(11))
    recon, mu, lv = model2(d["v"].to(DEV), c.to(DEV))
    kld = -0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp())
    bce = torch.nn.functional.binary_cross_entropy_with_logits(recon, d["tv"].to(DEV), reduction="sum")
    loss2 = (bce + klw * kld) / B
    loss2.backward()
    torch.cuda.synchronize()
    assert abs(loss2.item() - loss_o.item()) / loss_o.item() < 1e-4
    for k, p in model2.named_parameters():
        assert nrel(p.grad, sd_o[k].grad) < 1e-2, (k, nrel(p.grad, sd_o[k].grad))
    with pytest.raises(ValueError):
        model2(d["v"].to(DEV), c[:, :2].to(DEV))
    s = model2.inference(n=5, c=torch.rand(5, 3, device=DEV))
    assert s.shape == (5, 3, 64, 64) and torch.isfinite(s).all()


def test_cmvae_pose_fused_step_matches_oracle():
    """cnn-mvae --use-pose --conditional: 7 sub-sampled passes, every image encoder head / decoder sees
    the shock force, the pose expert does not (vae.py:118-123)."""
    from mmdyn_b200 import engine, noise
    B, klw, pm = 6, 0.02, 1000.0
    model, sd = make_cond("cnn-mvae", True, seed=1)
    d = batch(B, seed=4)
    c = 2 * torch.rand(B, 3, generator=torch.Generator().manual_seed(23)) - 1
    pkeys = [k for k, _ in model.named_parameters()]
    x_o, t_o = [d["v"], d["t"], d["p"]], [d["tv"], d["tt"], d["tp"]]
    sd_o = copy.deepcopy(sd)
    for k in pkeys:
        sd_o[k].requires_grad_(True)
    out_o, loss_o, per_pass = orc.evaluate_mvae(sd_o, x_o, t_o, klw, pm, True,
                                                oracle_noises(orc.MVAE_PASSES_POSE, B, 7), condition=c)
    loss_o.backward()
    eng = engine.StepEngine(model, "mvae", use_pose=True, pose_multiplier=pm,
                            noise_src=noise.HostNoise(torch.Generator().manual_seed(7)), exact_running_stats=True)
    out_d, loss_d = eng.evaluate([t.to(DEV) for t in x_o], [t.to(DEV) for t in t_o], klw, condition=c.to(DEV))
    torch.cuda.synchronize()
    assert abs(loss_d.item() - loss_o.item()) / loss_o.item() < 1e-4
    mu_d, lv_d = eng.ws.bufs["mu"], eng.ws.bufs["lv"]
    for i, pp in enumerate(per_pass):
        assert nrel(mu_d[i], pp["mu"]) < 3e-3 and nrel(lv_d[i], pp["lv"]) < 3e-3, i
    for a, b in zip(out_d["recon_x"], out_o["recon_x"]):
        assert nrel(a, b) < 2e-3
    loss_d.backward()
    torch.cuda.synchronize()
    gerr = {k: nrel(p.grad, sd_o[k].grad) for k, p in model.named_parameters()}
    worst = sorted(gerr.items(), key=lambda kv: -kv[1])[:5]
    print("worst gradient errors:\n" + "\n".join(f"{k:50s} {v:.3e}" for k, v in worst))
    for k, v in gerr.items():
        assert v < 5e-2, (k, v)  # same bound as the un-conditional pose case (ReLU flips behind z)
    flat_o = torch.cat([sd_o[k].grad.reshape(-1) for k in pkeys])
    flat_d = torch.cat([p.grad.reshape(-1) for _, p in model.named_parameters()])
    assert nrel(flat_d, flat_o) < 2e-2


# ---------------------------------------------------------------------------------------------
# Regressor baseline (models.py:28-77, problems.py:321-332)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cd", [0, 3])
def test_regressor_step_matches_oracle(cd):
    from mmdyn_b200 import losses, noise, optim
    from mmdyn_b200.pytorch.models.models import setup_model
    B = 8
    torch.manual_seed(3)
    model = setup_model("regressor", condition_dim=cd, out_dim=7, conditional=cd > 0)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(DEV)
    assert tuple(sd["out_net.0.weight"].shape) == (256, 512 + cd)
    d = batch(B, seed=12)
    c = (2 * torch.rand(B, 3, generator=torch.Generator().manual_seed(5)) - 1) if cd else None
    pkeys = [k for k, _ in model.named_parameters()]
    mask, _, _ = orc.draw_pass_noise(B, True, False, generator=torch.Generator().manual_seed(11))
    sd_o = copy.deepcopy(sd)
    for k in pkeys:
        sd_o[k].requires_grad_(True)
    acts = {}
    out_o, loss_o = orc.evaluate_regression(sd_o, d["v"], d["tp"], mask, c, acts=acts)
    loss_o.backward()
    model.noise = noise.HostNoise(torch.Generator().manual_seed(11))
    opt = optim.FusedAdam(model, lr=1e-3)
    opt.zero_grad()
    out = model(d["v"].to(DEV), c.to(DEV)) if cd else model(d["v"].to(DEV))
    loss = losses.mse_sum(out.view(B, 7), d["tp"].to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    # out = O(0.1) numbers behind a 12-layer fp16 trunk: norm-wise 2e-3; the loss is dominated by the targets
    assert nrel(out, out_o["outputs"]) < 2e-3, nrel(out, out_o["outputs"])
    assert abs(loss.item() - loss_o.item()) / loss_o.item() < 1e-3
    gerr = {k: nrel(p.grad, sd_o[k].grad) for k, p in model.named_parameters()}
    print("worst gradient errors:\n" + "\n".join(f"{k:40s} {v:.3e}" for k, v in sorted(gerr.items(), key=lambda kv: -kv[1])[:6]))
    for k, v in gerr.items():
        assert v < 2e-2, (k, v)  # ReLU units of out_net flip under the trunk's 1e-3 perturbation (cf. pose expert)
    flat_o = torch.cat([sd_o[k].grad.reshape(-1) for k in pkeys])
    flat_d = torch.cat([p.grad.reshape(-1) for _, p in model.named_parameters()])
    assert nrel(flat_d, flat_o) < 1e-2, nrel(flat_d, flat_o)
    if cd:
        e = nrel(dict(model.named_parameters())["out_net.0.weight"].grad[:, 512:], sd_o["out_net.0.weight"].grad[:, 512:])
        assert e < 2e-2, e
        with pytest.raises(ValueError):
            model(d["v"].to(DEV))
    # BatchNorm buffers follow the reference (one forward = one update)
    sdd = model.state_dict()
    for k in sd_o:
        if k.endswith("num_batches_tracked"):
            assert int(sdd[k]) == int(sd_o[k]) == 1, k
        elif "running_" in k:
            assert nrel(sdd[k], sd_o[k]) < 2e-3, k
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    opt.step()
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        assert torch.isfinite(p).all() and (p - before[k]).abs().max().item() <= 1.0001e-3, k


# ---------------------------------------------------------------------------------------------
# BASELINE.json's full sizes through a size-independent property: replication invariance
# ---------------------------------------------------------------------------------------------
class _ReplayNoise:
    """Test-only noise source handing out prescribed tensors in the engine's draw order."""

    def __init__(self, noises, passes, reps):
        self.q = []
        for (mv, mt, eps), (hv, ht, hp) in zip(noises, passes):
            self.q += ([("mask", mv)] if hv else []) + ([("mask", mt)] if ht else []) + [("eps", eps)]
        self.reps = reps

    def _next(self, kind, device, out):
        k, t = self.q.pop(0)
        assert k == kind
        t = t.repeat(self.reps, 1).to(device)
        if out is None:
            return t
        out.copy_(t)
        return out

    def dropout_mask(self, B, device, out=None):
        return self._next("mask", device, out)

    def normal(self, B, D, device, out=None):
        return self._next("eps", device, out)


@pytest.mark.parametrize("B", [1024, 4096])
def test_full_size_step_equals_oracle_on_the_replicated_base_batch(B):
    """The bench workload (cnn-mvae visuotactile + pose, 7 passes, per-GPU batch 1024; 4096 = the top of
    configs[4]'s sweep) cannot be run by the CPU oracle in seconds, but a batch made of R copies of a
    B0-row base batch (inputs, targets, dropout masks and eps all replicated) has the SAME BatchNorm
    batch statistics (mean and biased variance), the same per-row activations, the same loss (sum / B)
    and the same parameter gradients as the base batch.  So the full-size GPU step is checked against
    the oracle run on B0 = 16 rows: every tile, group and row offset of the big launch is exercised and
    any row-dependent indexing error breaks the equality."""
    from mmdyn_b200 import engine
    B0 = 16
    R = B // B0
    model, sd = make("cnn-mvae", True, seed=6)
    d = batch(B0, seed=14)
    klw, pm = 0.02, 1000.0
    passes = orc.MVAE_PASSES_POSE
    noises = oracle_noises(passes, B0, 31)
    pkeys = [k for k, _ in model.named_parameters()]
    x_o, t_o = [d["v"], d["t"], d["p"]], [d["tv"], d["tt"], d["tp"]]
    sd_o = copy.deepcopy(sd)
    for k in pkeys:
        sd_o[k].requires_grad_(True)
    out_o, loss_o, per_pass = orc.evaluate_mvae(sd_o, x_o, t_o, klw, pm, True, noises)
    loss_o.backward()
    rep = lambda t: t.repeat(R, *([1] * (t.dim() - 1))).to(DEV)
    eng = engine.StepEngine(model, "mvae", use_pose=True, pose_multiplier=pm, noise_src=_ReplayNoise(noises, passes, R))
    out_d, loss_d = eng.evaluate([rep(t) for t in x_o], [rep(t) for t in t_o], klw)
    loss_d.backward()
    torch.cuda.synchronize()
    assert abs(loss_d.item() - loss_o.item()) / loss_o.item() < 1e-4, (loss_d.item(), loss_o.item())
    for k, v in out_o["perf_measure"].items():
        assert abs(float(out_d["perf_measure"][k]) - v) / abs(v) < 1e-3, k
    mu_d, lv_d = eng.ws.bufs["mu"], eng.ws.bufs["lv"]          # (7, B, 256)
    for i, pp in enumerate(per_pass):
        assert nrel(mu_d[i][:B0], pp["mu"]) < 3e-3 and nrel(lv_d[i][:B0], pp["lv"]) < 3e-3, i
    # every replica carries the same rows: posterior and reconstructions agree across the whole batch.
    # Not bit for bit: the split-K fp32 atomics of the fc layer order their partial sums differently from
    # tile to tile (1e-7), and behind that a few fp16 roundings of the decoder chain flip by one ulp.
    for t in (mu_d, lv_d):
        blocks = t.view(t.shape[0], R, B0, -1)
        # (the fp16 copy of the fc feature that feeds the heads flips a rounding here and there: measured
        #  5e-6 .. 1.1e-5 norm-wise from run to run, bound 1e-4)
        assert nrel(blocks[:, R - 1], blocks[:, 0]) < 1e-4 and nrel(blocks[:, R // 2], blocks[:, 0]) < 1e-4
    for a, b in zip(out_d["recon_x"], out_o["recon_x"]):
        a = a.reshape(R, B0, -1)
        assert nrel(a[0], b.reshape(B0, -1)) < 3e-3
        assert nrel(a[R - 1], a[0]) < 1e-3 and nrel(a[R // 3], a[0]) < 1e-3
    gerr = {k: nrel(p.grad, sd_o[k].grad) for k, p in model.named_parameters()}
    worst = sorted(gerr.items(), key=lambda kv: -kv[1])[:5]
    print(f"B={B}: worst gradient errors vs the oracle on the base batch:\n" +
          "\n".join(f"{k:50s} {v:.3e}" for k, v in worst))
    for k, v in gerr.items():
        assert v < 5e-2, (k, v)
    flat_o = torch.cat([sd_o[k].grad.reshape(-1) for k in pkeys])
    flat_d = torch.cat([p.grad.reshape(-1) for _, p in model.named_parameters()])
    assert nrel(flat_d, flat_o) < 2e-2, nrel(flat_d, flat_o)
    # BatchNorm running statistics: the mean is replication invariant; the unbiased variance carries
    # n/(n-1) with n = rows of the big batch instead of the base batch
    sdd = model.state_dict()
    for k in sd_o:
        if k.endswith("running_mean") and "encoder" in k:
            assert nrel(sdd[k], sd_o[k]) < 3e-3, k


@pytest.mark.parametrize("kind", ["vae_tactile", "mvae_masked"])
def test_full_size_single_modality_and_masked_steps_equal_oracle_on_the_base_batch(kind):
    """Replication invariance (see above) for the other two BASELINE.json workloads at full size:
    configs[1] cnn-vae --input-type tactile (one pass, no prior expert) at batch 4096, and the
    --mask-loss cnn-mvae step without the pose expert (3 passes, mask multiplies logits and targets,
    problems.py:408-411, 445-447) at batch 2048."""
    from mmdyn_b200 import engine
    B0 = 16
    klw = 0.02
    if kind == "vae_tactile":
        B, passes = 4096, [(True, False, False)]
        model, sd = make("cnn-vae", seed=8)
    else:
        B, passes = 2048, orc.MVAE_PASSES_NOPOSE
        model, sd = make("cnn-mvae", False, seed=8)
    R = B // B0
    d = batch(B0, seed=15)
    gm = torch.Generator().manual_seed(3)
    mask = (torch.rand(B0, 3, 64, 64, generator=gm) > 0.5).float()
    noises = oracle_noises(passes, B0, 33)
    pkeys = [k for k, _ in model.named_parameters()]
    sd_o = copy.deepcopy(sd)
    for k in pkeys:
        sd_o[k].requires_grad_(True)
    rep = lambda t: t.repeat(R, *([1] * (t.dim() - 1))).to(DEV)
    if kind == "vae_tactile":
        mk, _, eps = noises[0]
        out_o, loss_o = orc.evaluate_vae(sd_o, d["t"], d["tt"], klw, (mk, eps), input_type="tactile")
        eng = engine.StepEngine(model, "vae", noise_src=_ReplayNoise(noises, passes, R))
        out_d, loss_d = eng.evaluate(rep(d["t"]), rep(d["tt"]), klw)
        rec_d, rec_o = [out_d["recon_x"]], [out_o["recon_x"]]
    else:
        out_o, loss_o, _ = orc.evaluate_mvae(sd_o, [d["v"], d["t"]], [d["tv"], d["tt"]], klw, 1000.0, False, noises,
                                             loss_mask=mask)
        eng = engine.StepEngine(model, "mvae", noise_src=_ReplayNoise(noises, passes, R))
        out_d, loss_d = eng.evaluate([rep(d["v"]), rep(d["t"])], [rep(d["tv"]), rep(d["tt"])], klw, loss_mask=rep(mask))
        rec_d, rec_o = out_d["recon_x"], out_o["recon_x"]
    loss_o.backward()
    loss_d.backward()
    torch.cuda.synchronize()
    assert abs(loss_d.item() - loss_o.item()) / loss_o.item() < 1e-4, (loss_d.item(), loss_o.item())
    assert nrel(out_d["means"][:B0], out_o["means"]) < 3e-3 and nrel(out_d["log_var"][:B0], out_o["log_var"]) < 3e-3
    for a, b in zip(rec_d, rec_o):
        a = a.reshape(R, B0, -1)
        assert nrel(a[0], b.reshape(B0, -1)) < 3e-3
        assert nrel(a[R - 1], a[0]) < 1e-3 and nrel(a[R // 3], a[0]) < 1e-3
    gerr = {k: nrel(p.grad, sd_o[k].grad) for k, p in model.named_parameters()}
    print(kind, "worst gradient errors:", sorted(gerr.items(), key=lambda kv: -kv[1])[:3])
    for k, v in gerr.items():
        assert v < 1e-2, (k, v)
    flat_o = torch.cat([sd_o[k].grad.reshape(-1) for k in pkeys])
    flat_d = torch.cat([p.grad.reshape(-1) for _, p in model.named_parameters()])
    assert nrel(flat_d, flat_o) < 3e-3, nrel(flat_d, flat_o)


def test_scaled_loss_backward_fails_loudly():
    """The fused backward is the gradient of `loss` for an upstream gradient of 1: (2 * loss).backward() must
    raise instead of silently producing unscaled gradients."""
    from mmdyn_b200 import engine, noise
    model, _ = make("cnn-vae", seed=1)
    d = batch(4, seed=2)
    eng = engine.StepEngine(model, "vae", noise_src=noise.HostNoise(torch.Generator().manual_seed(1)))
    _, loss = eng.evaluate(d["v"].to(DEV), d["tv"].to(DEV), 0.02)
    with pytest.raises(RuntimeError, match="upstream"):
        (2.0 * loss).backward()
    _, loss = eng.evaluate(d["v"].to(DEV), d["tv"].to(DEV), 0.02)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
