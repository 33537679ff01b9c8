"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported live from
/root/reference) on seeded synthetic, dataset-shaped inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference has no tests / golden vectors for mmdyn/pytorch, so these fixtures are
the pin for `oracle/mmdyn_oracle.py` (tests/test_oracle_cpu.py) and, through it, for the CUDA path.

Recipe: stub the two utility modules that fail headless (`utils.training` runs
`stty size` at import; `utils.plots` needs pyquaternion/matplotlib), build a dataset-free
SeqModeling/DynModeling object with object.__new__, call the reference's own
_evaluate_model / parse_input / optimizer.  Weights come from torch.manual_seed(seed) + the
reference's setup_model; fixtures store small tensors plus per-tensor checksums (sum, abs-sum,
first 8 elements) for the big ones, so the files stay a few hundred KB.
"""
import os
import sys
import types

import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def import_reference():
    sys.path.insert(0, REF)
    tr = types.ModuleType("mmdyn.pytorch.utils.training")
    tr.progress_bar = lambda *a, **k: None
    tr.save_pkl = lambda *a, **k: None
    pl = types.ModuleType("mmdyn.pytorch.utils.plots")
    pl.plot_pose_tensorboard = pl.plot_single_pose_tensorboard = lambda *a, **k: None
    sys.modules["mmdyn.pytorch.utils.training"] = tr
    sys.modules["mmdyn.pytorch.utils.plots"] = pl
    from mmdyn.pytorch.problems import problems
    from mmdyn.pytorch.models.models import setup_model
    return problems, setup_model


def summary(t):
    t = t.detach().double().reshape(-1)
    return {"sum": t.sum().item(), "abs": t.abs().sum().item(), "sq": t.pow(2).sum().item(),
            "head": t[:8].float().clone(), "numel": t.numel()}


def make_problem(problems, setup_model, cls, model_name, input_type, use_pose, seed, kl_weight, mask_loss=False,
                 cond_dim=0):
    pr = object.__new__(cls)
    pr.parameters = {"model_name": model_name, "input_type": input_type, "use_pose": use_pose,
                     "mask_loss": mask_loss, "problem_type": "seq_modeling"}
    pr._kl_weight, pr._pose_multiplier, pr._conditional = kl_weight, 1000.0, cond_dim > 0
    pr._device = torch.device("cpu")
    torch.manual_seed(seed)
    kw = dict(condition_dim=cond_dim, input_dim=4096, architecture="cnn", conditional=cond_dim > 0,
              categorical_conditions=False, latent_size=256)
    if "mvae" in model_name:
        kw["use_pose"] = use_pose
    pr._model = setup_model(model_name, cross_modal=input_type == "visuotactile", **kw)
    pr._model.train()
    return pr


def batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    d = dict(v=r(B, 3, 64, 64), t=r(B, 3, 64, 64), p=r(B, 7), tv=r(B, 3, 64, 64), tt=r(B, 3, 64, 64), tp=r(B, 7),
             mask=(r(B, 3, 64, 64) > 0.5).float())
    d["c"] = 2 * r(B, 3) - 1  # shock force (drawn last, so the fields above are what they always were)
    return d


def run_step(pr, x, targets, noise_seed):
    """problems.py:150-155 on the reference objects."""
    opt = torch.optim.Adam(pr._model.parameters(), lr=1e-3)
    opt.zero_grad()
    torch.manual_seed(noise_seed)  # the reference draws dropout masks + eps from the default generator
    outputs, loss = pr._evaluate_model(x, targets)
    loss.backward()
    names = [n for n, _ in pr._model.named_parameters()]
    grads = {n: p.grad.detach().clone() for n, p in pr._model.named_parameters()}
    opt.step()
    return outputs, loss.detach(), names, grads


def golden_case(problems, setup_model, name, model_name, input_type, use_pose, B, mask_loss=False, cond_dim=0):
    """cond_dim > 0: --conditional (CVAE), the shock force d["c"] is the condition (problems.py:683-703)."""
    pr = make_problem(problems, setup_model, problems.SeqModeling, model_name, input_type, use_pose, seed=0,
                      kl_weight=1.0 / 50, mask_loss=mask_loss, cond_dim=cond_dim)
    d = batch(B, seed=1)
    if input_type == "visuotactile":
        x = {"model_input": [d["v"], d["t"]], "input_object_pose": [d["p"]], "shock": None}
        t = {"target_output": [d["tv"], d["tt"]], "target_object_pose": [d["tp"]], "loss_mask": d["mask"]}
    else:
        k = "v" if input_type == "visual" else "t"
        x = {"model_input": d[k], "input_object_pose": None, "shock": None}
        t = {"target_output": d["t" + k], "target_object_pose": None, "loss_mask": d["mask"]}
    if cond_dim:
        assert cond_dim == d["c"].shape[1]
        x["shock"] = d["c"]
    w0 = {n: summary(p) for n, p in pr._model.state_dict().items() if p.is_floating_point()}
    outputs, loss, names, grads = run_step(pr, x, t, noise_seed=123)
    rec = outputs["recon_x"]
    rec = rec if isinstance(rec, (list, tuple)) else [rec]
    g = {"case": name, "model_name": model_name, "input_type": input_type, "use_pose": use_pose, "B": B,
         "mask_loss": mask_loss, "cond_dim": cond_dim, "weights_seed": 0, "data_seed": 1, "noise_seed": 123, "kl_weight": 1.0 / 50,
         "loss": loss.item(), "means": outputs["means"].detach().clone(), "log_var": outputs["log_var"].detach().clone(),
         "perf_measure": dict(outputs["perf_measure"]),
         "recon_x": [summary(r) for r in rec], "recon_head": [r.detach().reshape(-1)[:64].clone() for r in rec],
         "w0": w0, "grads": {n: summary(v) for n, v in grads.items()},
         "params_after": {n: summary(p) for n, p in pr._model.named_parameters()},
         "buffers_after": {n: b.detach().clone() for n, b in pr._model.named_buffers()
                           if n.endswith("num_batches_tracked") or b.numel() <= 64},
         "torch": torch.__version__, "threads": torch.get_num_threads()}
    torch.save(g, os.path.join(OUT, name + ".pt"))
    print(name, "loss", g["loss"], "perf", g["perf_measure"])


def golden_regression(problems, setup_model, name, B, cond_dim=0):
    """Regression._evaluate_model + Adam on the reference's Regressor (models.py:28-77, problems.py:263-332).
    The reference's own set_model cannot build it (it passes `condition_dim`, Regressor takes `num_classes`:
    TypeError), so the model is built through setup_model with the keywords Regressor does accept."""
    pr = object.__new__(problems.Regression)
    pr.parameters = {"model_name": "regressor", "input_type": "visual"}
    pr._conditional, pr._device, pr._seq_length = cond_dim > 0, torch.device("cpu"), 5
    torch.manual_seed(0)
    pr._model = setup_model("regressor", out_dim=7, conditional=cond_dim > 0, num_classes=cond_dim)  # num_classes=None fails even un-conditional: False * None (models.py:36)
    pr._model.train()
    pr.set_criterion()
    d = batch(B, seed=1)
    w0 = {n: summary(p) for n, p in pr._model.state_dict().items() if p.is_floating_point()}
    opt = torch.optim.Adam(pr._model.parameters(), lr=1e-3)
    opt.zero_grad()
    torch.manual_seed(123)
    outputs, loss = pr._evaluate_model({"model_input": d["v"], "shock": d["c"] if cond_dim else None}, d["tp"])
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in pr._model.named_parameters()}
    opt.step()
    # parse_input on a seq-collated batch (S sequences x L frames, small images)
    S, L = 3, 5
    g = torch.Generator().manual_seed(5)
    n = S * L
    data = [torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 7, generator=g),
            torch.ones(n, 2), torch.rand(n, 3, generator=g)]
    target = [torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 7, generator=g),
              (torch.rand(n, 3, 4, 4, generator=g) > 0.5).float()]
    parsed = {}
    for it in ("visual", "tactile"):
        pr.parameters["input_type"] = it
        parsed[it] = pr.parse_input([t.clone() for t in data], [t.clone() for t in target])
    out = {"case": name, "B": B, "cond_dim": cond_dim, "weights_seed": 0, "data_seed": 1, "noise_seed": 123,
           "loss": loss.item(), "outputs": outputs["outputs"].detach().clone(),
           "perf_measure": {k: float(v) for k, v in outputs["perf_measure"].items()},
           "w0": w0, "grads": {n: summary(v) for n, v in grads.items()},
           "params_after": {n: summary(p) for n, p in pr._model.named_parameters()},
           "buffers_after": {n: b.detach().clone() for n, b in pr._model.named_buffers()
                             if n.endswith("num_batches_tracked") or b.numel() <= 64},
           "parse": {"L": L, "data": data, "target": target, "parsed": parsed},
           "state_keys": list(pr._model.state_dict().keys()),
           "torch": torch.__version__, "threads": torch.get_num_threads()}
    torch.save(out, os.path.join(OUT, name + ".pt"))
    print(name, "loss", out["loss"], "perf", out["perf_measure"])


def golden_losses(problems):
    """Reconstruction._elbo_loss / _mvae_elbo_loss called directly (problems.py:401-458) on seeded tensors:
    the scalar training form (reduce=None), the per-sample scoring form (reduce=False) and the --mask-loss
    variants.  No model involved; pins oracle.elbo_loss / mvae_elbo_loss / *_per_sample."""
    import warnings
    pr = object.__new__(problems.Reconstruction)
    pr._kl_weight, pr._pose_multiplier = 0.3, 1000.0
    g = torch.Generator().manual_seed(11)
    B = 5
    rv, rt = torch.randn(B, 3, 64, 64, generator=g), torch.randn(B, 3, 64, 64, generator=g)
    tv, tt = torch.rand(B, 3, 64, 64, generator=g), torch.rand(B, 3, 64, 64, generator=g)
    rp, tp = torch.rand(B, 7, generator=g), torch.rand(B, 7, generator=g)
    mu, lv = torch.randn(B, 256, generator=g), 0.3 * torch.randn(B, 256, generator=g)
    mask = (torch.rand(B, 3, 64, 64, generator=g) > 0.5).float()
    out = {"seed": 11, "B": B, "kl_weight": 0.3, "pose_multiplier": 1000.0}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # the reference passes the deprecated `reduce=` argument
        out["elbo"] = pr._elbo_loss(rv, tv, mu, lv)
        out["elbo_masked"] = pr._elbo_loss(rv, tv, mu, lv, loss_mask=mask)
        out["elbo_ps"] = pr._elbo_loss(rv, tv, mu, lv, reduce=False)
        out["elbo_ps_masked"] = pr._elbo_loss(rv, tv, mu, lv, loss_mask=mask, reduce=False)
        out["mvae"] = pr._mvae_elbo_loss([rv, rt, rp], [tv, tt, tp], mu, lv)
        out["mvae_masked"] = pr._mvae_elbo_loss([rv, rt], [tv, tt], mu, lv, loss_mask=mask)
        out["mvae_ps"] = pr._mvae_elbo_loss([rv, rt, rp], [tv, tt, tp], mu, lv, reduce=False)
        out["mvae_ps_masked"] = pr._mvae_elbo_loss([rv, rt], [tv, tt], mu, lv, loss_mask=mask, reduce=False)
    torch.save({k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in out.items()},
               os.path.join(OUT, "losses.pt"))
    print("loss fixtures written:", {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})


def golden_parse_input(problems):
    """Integer / index path: SeqModeling.parse_input and DynModeling.parse_input on a seq-collated
    batch of S sequences x L frames (small images: the indexing does not depend on H, W)."""
    S, L = 3, 5
    g = torch.Generator().manual_seed(5)
    n = S * L
    data = [torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 7, generator=g),
            torch.ones(n, 2)]
    target = [torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 7, generator=g),
              (torch.rand(n, 3, 4, 4, generator=g) > 0.5).float()]
    out = {"S": S, "L": L, "data": data, "target": target}
    for cls, key in ((problems.SeqModeling, "seq"), (problems.DynModeling, "dyn")):
        for it in ("visual", "tactile", "visuotactile"):
            pr = object.__new__(cls)
            pr.parameters = {"input_type": it}
            pr._seq_length, pr._device = L, torch.device("cpu")
            xi, ti = pr.parse_input([t.clone() for t in data], [t.clone() for t in target])
            out[f"{key}.{it}"] = (xi, ti)
    torch.save(out, os.path.join(OUT, "parse_input.pt"))
    print("parse_input fixtures written")


def golden_anneal(problems):
    pr = object.__new__(problems.Problem)
    vals = []
    for ae in (50, 3):
        pr.parameters = {"annealing_epochs": ae}
        for e in range(6):
            pr._anneal_KL(e)
            vals.append((ae, e, pr._kl_weight))
    torch.save(vals, os.path.join(OUT, "anneal_kl.pt"))


if __name__ == "__main__":
    problems, setup_model = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "conditional":  # only the CVAE fixtures
        golden_case(problems, setup_model, "cvae_visual_b4", "cnn-vae", "visual", False, 4, cond_dim=3)
        golden_case(problems, setup_model, "cmvae_pose_b3", "cnn-mvae", "visuotactile", True, 3, cond_dim=3)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "losses":
        golden_losses(problems)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "regression":  # only the Regressor fixtures
        golden_regression(problems, setup_model, "regressor_b4", 4)
        golden_regression(problems, setup_model, "regressor_cond_b4", 4, cond_dim=3)
        sys.exit(0)
    golden_case(problems, setup_model, "vae_visual_b4", "cnn-vae", "visual", False, 4)
    golden_case(problems, setup_model, "vae_tactile_masked_b4", "cnn-vae", "tactile", False, 4, mask_loss=True)
    golden_case(problems, setup_model, "mvae_b4", "cnn-mvae", "visuotactile", False, 4)
    golden_case(problems, setup_model, "mvae_pose_b4", "cnn-mvae", "visuotactile", True, 4)
    golden_case(problems, setup_model, "mvae_masked_b3", "cnn-mvae", "visuotactile", False, 3, mask_loss=True)
    golden_case(problems, setup_model, "cvae_visual_b4", "cnn-vae", "visual", False, 4, cond_dim=3)
    golden_case(problems, setup_model, "cmvae_pose_b3", "cnn-mvae", "visuotactile", True, 3, cond_dim=3)
    golden_regression(problems, setup_model, "regressor_b4", 4)
    golden_regression(problems, setup_model, "regressor_cond_b4", 4, cond_dim=3)
    golden_losses(problems)
    golden_parse_input(problems)
    golden_anneal(problems)
