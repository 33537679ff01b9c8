"""Round-2 parity tests (VERDICT r1 "next" #1): the inference / sampling path with numbers (not shapes),
the forward-only validation step, a multi-step Adam trajectory and a one-step SGD update at the
reference's real batch size (B = 128, main.py:25), the device-side non-finite flag, and the two
advisor findings about CUDA graphs (stale packed weights after a replay, workspace buffers of other
batch sizes).  Everything is compared with the oracle (oracle/mmdyn_oracle.py) on identical weights,
inputs and noise; the measured errors are printed so the bounds can be read off the test log
(`pytest -s`, profiles/r2_parity_measured.txt)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mmdyn_oracle as orc  # noqa: E402

from .test_model_gpu import DEV, KW, _dataset_free_problem, batch, make, make_cond, nrel, oracle_noises  # noqa: E402


# ---------------------------------------------------------------------------------------------
# P16: MVAE.inference / VAE.inference (vae.py:90-98, 167-176), Reconstruction._sample + apply_sigmoid
# (problems.py:548-559, 616-626), _test_epoch's no_grad step (problems.py:173-191)
# ---------------------------------------------------------------------------------------------
def test_mvae_inference_and_sample_match_oracle():
    from mmdyn_b200 import noise
    n = 50  # Problem.train calls _sample(n=50) (problems.py:199)
    model, sd = make("cnn-mvae", True, seed=2)
    model.noise = noise.HostNoise(torch.Generator().manual_seed(5))
    z = torch.randn([n, 256], generator=torch.Generator().manual_seed(5))  # vae.py:172: the CPU draw, then .to(device)
    sd_o = copy.deepcopy(sd)
    v_o = orc.decoder_cnn(sd_o, "visual_decoder", z, track=True)
    t_o = orc.decoder_cnn(sd_o, "tactile_decoder", z, track=True)
    v, t = model.inference(n=n)
    torch.cuda.synchronize()
    ev, et = nrel(v, v_o), nrel(t, t_o)
    print(f"mvae.inference(n={n}): visual logits {ev:.3e}, tactile logits {et:.3e}")
    assert v.shape == (n, 3, 64, 64) and ev < 2e-3 and et < 2e-3
    # train-mode BatchNorm inside inference() tracks running statistics once per decoder, as the reference
    sdd = model.state_dict()
    for k in sd_o:
        if "decoder" in k and k.endswith("num_batches_tracked") and "pose" not in k:
            assert int(sdd[k]) == int(sd_o[k]) == 1, k
        elif "decoder" in k and "running_" in k:
            assert nrel(sdd[k], sd_o[k]) < 2e-3, k
    # _sample: sigmoid of the two reconstructions, stored for the image logger
    pr = _dataset_free_problem(model, True)
    pr._device, pr._categorical_conditions, pr._condition_dim = torch.device(DEV), False, 0
    from collections import defaultdict
    pr._img_logger_dict = defaultdict()
    model.noise = noise.HostNoise(torch.Generator().manual_seed(9))
    z2 = torch.randn([n, 256], generator=torch.Generator().manual_seed(9))
    pr._sample(n=n)
    got = pr._img_logger_dict['Samples/latent_space']
    want = [torch.sigmoid(orc.decoder_cnn(sd_o, p, z2, track=True)) for p in ("visual_decoder", "tactile_decoder")]
    for g_, w_ in zip(got, want):
        e = (g_.cpu() - w_).abs().max().item()
        print(f"_sample sigmoid: max abs err {e:.3e}")
        assert g_.min() >= 0 and g_.max() <= 1 and e < 2e-3


def test_vae_and_cvae_inference_match_oracle():
    from mmdyn_b200 import noise
    n = 12
    model, sd = make("cnn-vae", seed=4)
    model.noise = noise.HostNoise(torch.Generator().manual_seed(3))
    z = torch.randn([n, 256], generator=torch.Generator().manual_seed(3))
    want = orc.decoder_cnn(copy.deepcopy(sd), "decoder", z, track=True)
    got = model.inference(n=n)
    e = nrel(got, want)
    print(f"vae.inference: {e:.3e}")
    assert e < 2e-3
    # CVAE: the condition enters the upsample Linear (vae.py:286-291)
    cm, csd = make_cond("cnn-vae", seed=6)
    cm.noise = noise.HostNoise(torch.Generator().manual_seed(8))
    z = torch.randn([n, 256], generator=torch.Generator().manual_seed(8))
    c = torch.rand(n, 3, generator=torch.Generator().manual_seed(1))
    want = orc.decoder_cnn(copy.deepcopy(csd), "decoder", z, track=True, c=c)
    got = cm.inference(n=n, c=c.to(DEV))
    e = nrel(got, want)
    print(f"cvae.inference: {e:.3e}")
    assert e < 2e-3
    other = orc.decoder_cnn(copy.deepcopy(csd), "decoder", z, track=False, c=torch.zeros_like(c))
    assert nrel(other, want) > 1e-3  # the condition matters


@pytest.mark.parametrize("use_pose", [False, True])
def test_no_grad_validation_step_matches_oracle(use_pose):
    """Problem._test_epoch: `with torch.no_grad(): outputs, loss = self._evaluate_model(...)` in TRAIN mode
    (problems.py:174-182): forward-only fused step, no backward state kept, no gradient touched."""
    from mmdyn_b200 import noise
    B, klw = 8, 0.3
    model, sd = make("cnn-mvae", use_pose, seed=3)
    d = batch(B, seed=5)
    passes = orc.MVAE_PASSES_POSE if use_pose else orc.MVAE_PASSES_NOPOSE
    x_o = [d["v"], d["t"]] + ([d["p"]] if use_pose else [])
    t_o = [d["tv"], d["tt"]] + ([d["tp"]] if use_pose else [])
    with torch.no_grad():
        out_o, loss_o, _ = orc.evaluate_mvae(copy.deepcopy(sd), x_o, t_o, klw, 1000.0, use_pose, oracle_noises(passes, B, 13))
    pr = _dataset_free_problem(model, use_pose, klw)
    model.noise = noise.HostNoise(torch.Generator().manual_seed(13))
    inputs = {"model_input": [d["v"].to(DEV), d["t"].to(DEV)], "input_object_pose": [d["p"].to(DEV)]}
    targets = {"target_output": [d["tv"].to(DEV), d["tt"].to(DEV)], "target_object_pose": [d["tp"].to(DEV)],
               "loss_mask": None}
    model.train()
    with torch.no_grad():
        out_d, loss_d = pr._evaluate_model(inputs, targets)
    torch.cuda.synchronize()
    el = abs(loss_d.item() - loss_o.item()) / abs(loss_o.item())
    print(f"no_grad step (pose={use_pose}): loss rel {el:.3e}")
    assert el < 1e-4 and not loss_d.requires_grad and pr._engine._state is None
    for a, b in zip(out_d["recon_x"], out_o["recon_x"]):
        assert nrel(a, b) < 2e-3
    assert nrel(out_d["means"], out_o["means"]) < 2e-3 and nrel(out_d["log_var"], out_o["log_var"]) < 2e-3
    for k, v in out_o["perf_measure"].items():
        assert abs(float(out_d["perf_measure"][k]) - v) / abs(v) < 1e-3, k
    assert all(p.grad is None or float(p.grad.abs().sum()) == 0.0 for p in model.parameters())


# ---------------------------------------------------------------------------------------------
# trajectory: N optimizer steps at the reference's batch size (config 3: B = 128)
# ---------------------------------------------------------------------------------------------
def _b128_batches(n, use_pose=True):
    out = []
    for i in range(n):
        d = batch(128, seed=40 + i)
        out.append(([d["v"], d["t"]] + ([d["p"]] if use_pose else []), [d["tv"], d["tt"]] + ([d["tp"]] if use_pose else [])))
    return out


def test_adam_trajectory_20_steps_b128_matches_oracle():
    """cnn-mvae visuotactile + pose, B = 128, 20 iterations of problems.py:150-155 (zero_grad, 7-pass step,
    backward, Adam lr 1e-3) on three alternating batches, KL weight annealed per step like epochs would:
    the loss curve follows the oracle's step by step and the accumulated parameter displacement agrees."""
    from mmdyn_b200 import engine, noise, optim
    steps, pm = 20, 1000.0
    model, sd = make("cnn-mvae", True, seed=11)
    pkeys = [k for k, _ in model.named_parameters()]
    data = _b128_batches(3)
    # ---- oracle trajectory ----
    sd_o = copy.deepcopy(sd)
    st = {"step": 0, "m": [torch.zeros_like(sd_o[k]) for k in pkeys], "v": [torch.zeros_like(sd_o[k]) for k in pkeys]}
    g_o = torch.Generator().manual_seed(77)
    loss_o = []
    for i in range(steps):
        x, t = data[i % 3]
        noises = [orc.draw_pass_noise(128, hv, ht, generator=g_o) for (hv, ht, hp) in orc.MVAE_PASSES_POSE]
        _, l, _ = orc.train_step(sd_o, pkeys, "mvae+pose", {"x": x, "targets": t}, (i + 1) / 50, pm, noises, st)
        loss_o.append(float(l))
    # ---- B200 trajectory ----
    eng = engine.StepEngine(model, "mvae", use_pose=True, pose_multiplier=pm,
                            noise_src=noise.HostNoise(torch.Generator().manual_seed(77)))
    opt = optim.FusedAdam(model, lr=1e-3)
    dev_data = [([a.to(DEV) for a in x], [a.to(DEV) for a in t]) for x, t in data]
    loss_d = []
    for i in range(steps):
        x, t = dev_data[i % 3]
        opt.zero_grad()
        _, l = eng.evaluate(x, t, (i + 1) / 50, want_outputs=False)
        l.backward()
        opt.step()
        loss_d.append(float(l.detach()))
    opt.check_finite(loss_d[-1])
    errs = [abs(a - b) / abs(b) for a, b in zip(loss_d, loss_o)]
    print("Adam trajectory, B=128, loss per step (B200 | oracle | rel):")
    for i, (a, b, e) in enumerate(zip(loss_d, loss_o, errs)):
        print(f"  step {i:2d}  {a:12.4f}  {b:12.4f}  {e:.2e}")
    assert loss_o[-1] < 0.9 * loss_o[0], "the oracle trajectory itself must be learning"
    assert max(errs) < 5e-4, max(errs)  # measured (r2): max 1.9e-4 at step 2, <= 5e-5 elsewhere
    # accumulated displacement: relative L2 distance of (theta_20 - theta_0) over the whole arena and per sub-network
    num = den = 0.0
    per = {}
    for k, p in model.named_parameters():
        dd, do = (p.detach().cpu() - sd[k]).double(), (sd_o[k].detach() - sd[k]).double()
        a, b = (dd - do).pow(2).sum().item(), do.pow(2).sum().item()
        num, den = num + a, den + b
        top = k.split(".")[0]
        per[top] = (per.get(top, (0, 0))[0] + a, per.get(top, (0, 0))[1] + b)
    print("displacement after 20 steps, rel L2: all %.3e | " % ((num / den) ** 0.5) +
          ", ".join(f"{k} {(a / b) ** 0.5:.2e}" for k, (a, b) in per.items()))
    # step 1 of Adam is lr*sign(g): entries whose gradient sits below the fp16 noise floor start in a random
    # direction on either side; over 20 steps the moments average that out
    assert (num / den) ** 0.5 < 0.06, (num / den) ** 0.5  # measured (r2): 3.0e-2 over the arena


@pytest.mark.parametrize("use_pose", [False, True])
def test_sgd_one_step_b128_matches_oracle(use_pose):
    """--optimizer SGD (problems.py:132-136: momentum 0.9, weight decay 5e-4) at B = 128: the first update is
    -lr * (g + wd * p), i.e. proportional to the gradient, so the parameter deltas hold a real bound."""
    from mmdyn_b200 import engine, noise, optim
    lr, wd, klw = 1e-3, 5e-4, 0.02
    model, sd = make("cnn-mvae", use_pose, seed=12)
    pkeys = [k for k, _ in model.named_parameters()]
    (x, t), = _b128_batches(1, use_pose)
    passes = orc.MVAE_PASSES_POSE if use_pose else orc.MVAE_PASSES_NOPOSE
    sd_o = copy.deepcopy(sd)
    for k in pkeys:
        sd_o[k].requires_grad_(True)
    _, loss_o, _ = orc.evaluate_mvae(sd_o, x, t, klw, 1000.0, use_pose, oracle_noises(passes, 128, 19))
    loss_o.backward()
    delta_o = {k: -lr * (sd_o[k].grad + wd * sd_o[k].detach()) for k in pkeys}
    eng = engine.StepEngine(model, "mvae", use_pose=use_pose, noise_src=noise.HostNoise(torch.Generator().manual_seed(19)))
    opt = optim.FusedSGD(model, lr=lr, momentum=0.9, weight_decay=wd)
    opt.zero_grad()
    _, loss_d = eng.evaluate([a.to(DEV) for a in x], [a.to(DEV) for a in t], klw, want_outputs=False)
    loss_d.backward()
    gerr = {k: nrel(p.grad, sd_o[k].grad) for k, p in model.named_parameters()}
    opt.step()
    opt.check_finite(float(loss_d))
    torch.cuda.synchronize()
    assert abs(loss_d.item() - loss_o.item()) / loss_o.item() < 1e-4
    derr = {k: nrel(p.detach().cpu() - sd[k], delta_o[k]) for k, p in model.named_parameters()}
    flat_d = torch.cat([(p.detach().cpu() - sd[k]).reshape(-1) for k, p in model.named_parameters()])
    flat_o = torch.cat([delta_o[k].reshape(-1) for k in pkeys])
    worst = sorted(derr.items(), key=lambda kv: -kv[1])[:6]
    print(f"SGD one step, B=128, pose={use_pose}: whole-arena delta rel {nrel(flat_d, flat_o):.3e}; worst tensors:\n" +
          "\n".join(f"  {k:50s} delta {v:.3e}  grad {gerr[k]:.3e}" for k, v in worst))
    # measured (r2): whole arena 4.2e-4 / 4.5e-4; worst tensor 2.4e-3 without, 9.7e-3 with the pose expert (ReLU flips
    # of the pose decoder behind z: test_pose_gradient_error_is_explained_by_relu_flips)
    assert nrel(flat_d, flat_o) < 2e-3, nrel(flat_d, flat_o)
    for k, v in derr.items():
        assert v < (2.5e-2 if use_pose else 5e-3), (k, v)


# ---------------------------------------------------------------------------------------------
# why gradients move by ~1 % with the pose expert: ReLU units of the pose decoder flip under the 1e-3
# perturbation of z that the fp16 image trunks cause
# ---------------------------------------------------------------------------------------------
def test_pose_gradient_error_is_explained_by_relu_flips():
    """Rows of the pose-decoder input z whose two ReLU masks are IDENTICAL on both sides carry a dz that agrees
    to the precision of z itself; the rows with a flipped unit carry the percent-level error (a unit with
    pre-activation ~0 changes nothing in the forward pass but switches one rank-1 term of dz on or off)."""
    from mmdyn_b200 import engine, noise
    B, klw, pm = 64, 0.02, 1000.0
    model, sd = make("cnn-mvae", True, seed=6)
    d = batch(B, seed=14)
    x_o, t_o = [d["v"], d["t"], d["p"]], [d["tv"], d["tt"], d["tp"]]
    with torch.no_grad():
        _, _, per_pass = orc.evaluate_mvae(copy.deepcopy(sd), x_o, t_o, klw, pm, True,
                                           oracle_noises(orc.MVAE_PASSES_POSE, B, 31))
    eng = engine.StepEngine(model, "mvae", use_pose=True, pose_multiplier=pm,
                            noise_src=noise.HostNoise(torch.Generator().manual_seed(31)))
    _, loss = eng.evaluate([a.to(DEV) for a in x_o], [a.to(DEV) for a in t_o], klw, want_outputs=False)
    loss.backward()
    torch.cuda.synchronize()
    ws = eng.ws.bufs
    # pose passes 3..6 -> groups 0..3 of the pose decoder launch; z of pass i from the oracle's posterior + eps
    g_n = torch.Generator().manual_seed(31)
    eps = [orc.draw_pass_noise(B, hv, ht, generator=g_n)[2] for (hv, ht, hp) in orc.MVAE_PASSES_POSE]
    W = {k: sd["pose_decoder.deconv_net.%d.%s" % (i, n)].double() for i in (0, 2, 4) for n in ("weight", "bias")
         for k in [(i, n)]}
    same_rows, flip_rows, e_same, e_flip = 0, 0, [], []
    for g, i in enumerate((3, 4, 5, 6)):
        z = (eps[i] * torch.exp(0.5 * per_pass[i]["lv"]) + per_pass[i]["mu"]).double().requires_grad_(True)
        a1 = torch.relu(z @ W[(0, "weight")].T + W[(0, "bias")])
        a2 = torch.relu(a1 @ W[(2, "weight")].T + W[(2, "bias")])
        rec = a2 @ W[(4, "weight")].T + W[(4, "bias")]
        (pm * (rec - t_o[2].double()).pow(2).sum() / B).backward()
        dz_o = z.grad * B  # the fused backward carries gradients times grad_scale = B
        sl = slice(g * B, (g + 1) * B)
        a1_d, a2_d, dz_d = ws["pdec.a1"][sl].cpu(), ws["pdec.a2"][sl].cpu(), ws["pdec.dz"][sl].cpu().double()
        flips = ((a1_d > 0) != (a1.detach() > 0)).sum(1) + ((a2_d > 0) != (a2.detach() > 0)).sum(1)
        for r in range(B):
            e = ((dz_d[r] - dz_o[r]).norm() / dz_o[r].norm()).item()
            if flips[r] == 0:
                same_rows += 1
                e_same.append(e)
            else:
                flip_rows += 1
                e_flip.append(e)
    med = lambda v: sorted(v)[len(v) // 2] if v else float("nan")
    print(f"pose decoder dz, {same_rows} rows with identical ReLU masks: median rel err {med(e_same):.2e}, max {max(e_same):.2e}; "
          f"{flip_rows} rows with >= 1 flipped unit: median {med(e_flip):.2e}, max {max(e_flip) if e_flip else 0:.2e}")
    assert same_rows > 0 and max(e_same) < 1e-2 and med(e_same) < 3e-3
    if e_flip:
        assert med(e_flip) > 2 * med(e_same)


# ---------------------------------------------------------------------------------------------
# device-side non-finite flag (fp16 operands carry no overflow guard of their own)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("optname", ["Adam", "SGD"])
def test_nonfinite_gradients_trip_the_flag_instead_of_training(optname):
    from mmdyn_b200 import engine, noise, optim
    model, sd = make("cnn-vae", seed=1)
    d = batch(8, seed=2)
    eng = engine.StepEngine(model, "vae", noise_src=noise.HostNoise(torch.Generator().manual_seed(1)))
    opt = optim.FusedAdam(model, lr=1e-3) if optname == "Adam" else optim.FusedSGD(model, lr=1e-3)
    # a healthy step first: flag stays clear
    opt.zero_grad()
    _, loss = eng.evaluate(d["v"].to(DEV), d["tv"].to(DEV), 0.02, want_outputs=False)
    loss.backward()
    opt.step()
    opt.check_finite(float(loss))
    before = torch.cat([p.detach().reshape(-1).clone() for p in model.parameters()])
    # inputs scaled by 1e6 overflow the fp16 activations of the first layers
    opt.zero_grad()
    _, loss = eng.evaluate(1e6 * d["v"].to(DEV), d["tv"].to(DEV), 0.02, want_outputs=False)
    loss.backward()
    opt.step()
    assert int(opt.nonfinite_flag()) != 0
    with pytest.raises(FloatingPointError, match="non-finite"):
        opt.check_finite(float(loss))
    after = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.isfinite(after).all(), "a non-finite gradient entry reached the parameters"
    n_moved = int((after != before).sum())
    print(f"{optname}: poisoned step moved {n_moved} of {after.numel()} parameters (entries with a finite gradient)")
    # the flag is sticky until read, then clear: the next healthy step passes
    opt.zero_grad()
    _, loss = eng.evaluate(d["v"].to(DEV), d["tv"].to(DEV), 0.02, want_outputs=False)
    loss.backward()
    opt.step()
    opt.check_finite(float(loss))


# ---------------------------------------------------------------------------------------------
# advisor findings, round 1
# ---------------------------------------------------------------------------------------------
def test_eager_calls_after_graph_replays_see_the_updated_weights():
    """GraphedTrainStep.run() replays an optimizer step that rewrites the parameter arena; the fp16 operand
    copies must be re-packed for the eager validation / sampling that follows (Problem._test_epoch)."""
    from mmdyn_b200 import engine, noise, optim
    B, klw = 8, 0.02
    model, _ = make("cnn-mvae", False, seed=3)
    d = batch(B, seed=4)
    x, t = [d["v"].to(DEV), d["t"].to(DEV)], [d["tv"].to(DEV), d["tt"].to(DEV)]
    eng = engine.StepEngine(model, "mvae", noise_src=noise.DeviceNoise(seed=5))
    opt = optim.FusedAdam(model, lr=1e-2)
    g = engine.GraphedTrainStep(eng, opt, x, t, klw)
    for _ in range(5):
        g.run()

    def eager_loss(force):
        eng.noise_src = noise.HostNoise(torch.Generator().manual_seed(3))
        if force:
            engine.get_execs(model, None)[1]["packer"].token = None
        with torch.no_grad():
            _, l = eng.evaluate(x, t, klw, want_outputs=False)
        return float(l)
    a, b = eager_loss(False), eager_loss(True)
    print(f"eager loss after 5 graphed steps: {a:.6f}; after a forced re-pack: {b:.6f}")
    assert abs(a - b) <= 1e-6 * abs(b), (a, b)
    # and the weights did move (lr 1e-2 x 5 steps), so a stale pack would have shown
    eng.noise_src = noise.DeviceNoise(seed=5)


def test_graph_survives_eager_steps_of_other_batch_sizes():
    """Captured graphs hold raw pointers into Workspace buffers: evaluating another batch size on the same
    engine must not recycle them."""
    from mmdyn_b200 import engine, noise, optim
    klw = 0.02
    model, _ = make("cnn-mvae", False, seed=3)
    d = batch(16, seed=4)
    dev = lambda n: ([d["v"][:n].to(DEV), d["t"][:n].to(DEV)], [d["tv"][:n].to(DEV), d["tt"][:n].to(DEV)])
    src = noise.DeviceNoise(seed=5)
    eng = engine.StepEngine(model, "mvae", noise_src=src)
    opt = optim.FusedAdam(model, lr=1e-3)
    x8, t8 = dev(8)
    g = engine.GraphedTrainStep(eng, opt, x8, t8, klw, split_optimizer=True)  # forward + backward only
    arena = engine.get_arena(model)
    def replay():
        src.ctr.zero_()  # same dropout masks / eps every time
        g.run()
        torch.cuda.synchronize()
        return float(g.loss), arena.grad.clone()
    loss_a, grad_a = replay()
    loss_a2, grad_a2 = replay()
    # run-to-run noise floor of one and the same graph: fp32 atomics (split-K, BatchNorm sums, weight gradients)
    # land in a different order, and behind them a few fp16 roundings flip
    e0 = nrel(grad_a2, grad_a)
    for n in (4, 16, 5):  # other shapes through the same engine / workspace, eagerly
        xs, ts = dev(n)
        _, l = eng.evaluate(xs, ts, klw, want_outputs=False)
        l.backward()
    torch.cuda.synchronize()
    loss_b, grad_b = replay()
    e = nrel(grad_b, grad_a)
    print(f"graph replay before / after eager steps of other sizes: loss {loss_a:.6f} / {loss_b:.6f}, grad rel {e:.2e} "
          f"(replay-to-replay noise floor {e0:.2e}, loss {abs(loss_a2 - loss_a) / abs(loss_a):.1e})")
    assert abs(loss_a - loss_b) <= 2e-6 * abs(loss_a)
    assert e < max(5 * e0, 1e-5), (e, e0)
