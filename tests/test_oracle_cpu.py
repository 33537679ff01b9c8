"""Pins the oracle (oracle/mmdyn_oracle.py) to the golden fixtures generated from the live
reference (tests/golden/make_golden.py), and the host-side index logic (parse_input, KL annealing)
of the product mirror to the same fixtures, bit-exactly.  CPU only."""
import glob
import os

import pytest
import torch

from oracle import mmdyn_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KW = dict(condition_dim=0, input_dim=4096, architecture="cnn", conditional=False, categorical_conditions=False,
          latent_size=256)


def summary(t):
    t = t.detach().double().reshape(-1)
    return {"sum": t.sum().item(), "abs": t.abs().sum().item(), "sq": t.pow(2).sum().item(), "head": t[:8].float()}


def close(a, b, rtol):
    return abs(a - b) <= rtol * max(abs(a), abs(b), 1e-30)


def check_summary(mine, gold, rtol, what):
    s = summary(mine)
    assert close(s["abs"], gold["abs"], rtol), (what, "abs", s["abs"], gold["abs"])
    assert close(s["sq"], gold["sq"], rtol), (what, "sq", s["sq"], gold["sq"])
    assert torch.allclose(s["head"], gold["head"], rtol=max(rtol, 1e-5) * 20, atol=1e-7), (what, s["head"], gold["head"])


def batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    d = dict(v=r(B, 3, 64, 64), t=r(B, 3, 64, 64), p=r(B, 7), tv=r(B, 3, 64, 64), tt=r(B, 3, 64, 64), tp=r(B, 7),
             mask=(r(B, 3, 64, 64) > 0.5).float())
    d["c"] = 2 * r(B, 3) - 1  # shock force of the conditional (CVAE) fixtures
    return d


CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLD, "*vae*.pt")))


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_step(case):
    g = torch.load(os.path.join(GOLD, case + ".pt"), weights_only=False)
    torch.set_num_threads(g["threads"])
    from mmdyn_b200.pytorch.models.models import setup_model
    torch.manual_seed(g["weights_seed"])
    kw = dict(KW)
    cd = g.get("cond_dim", 0)
    if cd:
        kw.update(condition_dim=cd, conditional=True)
    if "mvae" in g["model_name"]:
        kw["use_pose"] = g["use_pose"]
    model = setup_model(g["model_name"], cross_modal=g["input_type"] == "visuotactile", **kw)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    # the mirror's initialisation is the reference's, bit for bit
    for k, s in g["w0"].items():
        m = summary(sd[k])
        assert m["sum"] == s["sum"] and m["abs"] == s["abs"], k
    pkeys = [k for k, _ in model.named_parameters()]
    B, d = g["B"], batch(g["B"], g["data_seed"])
    mask = d["mask"] if g["mask_loss"] else None
    cond = d["c"] if cd else None
    torch.manual_seed(g["noise_seed"])
    if "mvae" in g["model_name"]:
        passes = orc.MVAE_PASSES_POSE if g["use_pose"] else orc.MVAE_PASSES_NOPOSE
        noises = [orc.draw_pass_noise(B, hv, ht) for (hv, ht, hp) in passes]
        b = {"x": [d["v"], d["t"]] + ([d["p"]] if g["use_pose"] else []),
             "targets": [d["tv"], d["tt"]] + ([d["tp"]] if g["use_pose"] else [])}
        problem = "mvae+pose" if g["use_pose"] else "mvae"
    else:
        k = "v" if g["input_type"] == "visual" else "t"
        m_, _, e_ = orc.draw_pass_noise(B, True, False)
        noises = [(m_, e_)]
        b = {"x": d[k], "target": d["t" + k]}
        problem = "vae"
    st = {"step": 0, "m": [torch.zeros_like(sd[k]) for k in pkeys], "v": [torch.zeros_like(sd[k]) for k in pkeys]}
    if problem == "vae":
        params = [sd[k].requires_grad_(True) for k in pkeys]
        outputs, loss = orc.evaluate_vae(sd, b["x"], b["target"], g["kl_weight"], noises[0], mask,
                                         input_type=g["input_type"], condition=cond)
        loss.backward()
        grads = [p.grad.detach().clone() for p in params]
        with torch.no_grad():
            for p in params:
                p.requires_grad_(False)
            orc.adam_step(params, grads, st, lr=1e-3)
        loss = loss.detach()
    else:
        outputs, loss, grads = orc.train_step(sd, pkeys, problem, b, g["kl_weight"], 1000.0, noises, st, lr=1e-3,
                                              loss_mask=mask, condition=cond)
    # identical op sequence on identical inputs: fp32 reduction-order noise only
    assert close(loss.item(), g["loss"], 2e-6), (loss.item(), g["loss"])
    assert torch.allclose(outputs["means"], g["means"], rtol=1e-4, atol=1e-6)
    assert torch.allclose(outputs["log_var"], g["log_var"], rtol=1e-4, atol=1e-6)
    for k, v in g["perf_measure"].items():
        assert close(outputs["perf_measure"][k], v, 1e-5), k
    rec = outputs["recon_x"] if isinstance(outputs["recon_x"], (list, tuple)) else [outputs["recon_x"]]
    for r, s, h in zip(rec, g["recon_x"], g["recon_head"]):
        check_summary(r, s, 1e-4, "recon")
        assert torch.allclose(r.detach().reshape(-1)[:64], h, rtol=1e-3, atol=1e-5)
    for k, gr in zip(pkeys, grads):
        check_summary(gr, g["grads"][k], 2e-3, "grad " + k)
    for k in pkeys:
        check_summary(sd[k], g["params_after"][k], 1e-4, "param " + k)
    for k, v in g["buffers_after"].items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v), k
        else:
            assert torch.allclose(sd[k], v, rtol=1e-4, atol=1e-6), k


def _mirror(cls_name, it, L):
    from mmdyn_b200.pytorch.problems import problems
    pr = object.__new__(getattr(problems, cls_name))
    pr.parameters = {"input_type": it}
    pr._seq_length, pr._device = L, torch.device("cpu")
    return pr


def _same(a, b):
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return torch.equal(a, b)


@pytest.mark.parametrize("key,cls", [("seq", "SeqModeling"), ("dyn", "DynModeling")])
@pytest.mark.parametrize("it", ["visual", "tactile", "visuotactile"])
def test_parse_input_is_bit_exact(key, cls, it):
    g = torch.load(os.path.join(GOLD, "parse_input.pt"), weights_only=False)
    xi_g, ti_g = g[f"{key}.{it}"]
    pr = _mirror(cls, it, g["L"])
    xi, ti = pr.parse_input([t.clone() for t in g["data"]], [t.clone() for t in g["target"]])
    fn = orc.seq_parse_input if key == "seq" else orc.dyn_parse_input
    xo, to = fn([t.clone() for t in g["data"]], [t.clone() for t in g["target"]], g["L"], it)
    for got_x, got_t in ((xi, ti), (xo, to)):
        for k in xi_g:
            assert _same(got_x[k], xi_g[k]), (key, it, k)
        for k in ti_g:
            assert _same(got_t[k], ti_g[k]), (key, it, k)
    if key == "dyn":
        # the reference's wrap-around quirk (problems.py:798): the last pose target is row 0's pose
        assert torch.equal(ti["target_object_pose"][0][-1], g["data"][2][0])


def test_anneal_kl_matches_reference():
    from mmdyn_b200.pytorch.problems import problems
    vals = torch.load(os.path.join(GOLD, "anneal_kl.pt"), weights_only=False)
    pr = object.__new__(problems.Problem)
    for ae, e, w in vals:
        pr.parameters = {"annealing_epochs": ae}
        pr._anneal_KL(e)
        assert pr._kl_weight == w == orc.anneal_kl(e, ae)


# ---------------------------------------------------------------------------------------------
# Regressor / --problem-type regression (models.py:28-77, problems.py:263-332)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["regressor_b4", "regressor_cond_b4"])
def test_oracle_reproduces_reference_regression_step(case):
    g = torch.load(os.path.join(GOLD, case + ".pt"), weights_only=False)
    torch.set_num_threads(g["threads"])
    from mmdyn_b200.pytorch.models.models import setup_model
    cd = g["cond_dim"]
    torch.manual_seed(g["weights_seed"])
    # the keywords the reference's Regression.set_model passes (problems.py:272-277)
    model = setup_model("regressor", condition_dim=cd, out_dim=7, conditional=cd > 0)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    assert list(sd.keys()) == g["state_keys"]
    for k, s in g["w0"].items():  # bit-identical initialisation
        m = summary(sd[k])
        assert m["sum"] == s["sum"] and m["abs"] == s["abs"], k
    pkeys = [k for k, _ in model.named_parameters()]
    d = batch(g["B"], g["data_seed"])
    torch.manual_seed(g["noise_seed"])
    mask, _, _ = orc.draw_pass_noise(g["B"], True, False)  # the Dropout(0.1) mask behind fc_net
    params = [sd[k].requires_grad_(True) for k in pkeys]
    outputs, loss = orc.evaluate_regression(sd, d["v"], d["tp"], mask, d["c"] if cd else None)
    loss.backward()
    grads = [p.grad.detach().clone() for p in params]
    st = {"step": 0, "m": [torch.zeros_like(p) for p in params], "v": [torch.zeros_like(p) for p in params]}
    with torch.no_grad():
        for p in params:
            p.requires_grad_(False)
        orc.adam_step(params, grads, st, lr=1e-3)
    assert close(loss.item(), g["loss"], 2e-6), (loss.item(), g["loss"])
    assert torch.allclose(outputs["outputs"].detach(), g["outputs"], rtol=1e-4, atol=1e-6)
    assert close(outputs["perf_measure"]["pose"], g["perf_measure"]["pose"], 1e-5)
    for k, gr in zip(pkeys, grads):
        check_summary(gr, g["grads"][k], 2e-3, "grad " + k)
    for k in pkeys:
        check_summary(sd[k], g["params_after"][k], 1e-4, "param " + k)
    for k, v in g["buffers_after"].items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v), k
        else:
            assert torch.allclose(sd[k], v, rtol=1e-4, atol=1e-6), k


@pytest.mark.parametrize("it", ["visual", "tactile"])
def test_regression_parse_input_is_bit_exact(it):
    g = torch.load(os.path.join(GOLD, "regressor_cond_b4.pt"), weights_only=False)["parse"]
    from mmdyn_b200.pytorch.problems import problems
    pr = object.__new__(problems.Regression)
    pr.parameters = {"input_type": it}
    pr._seq_length, pr._device = g["L"], torch.device("cpu")
    xi, ti = pr.parse_input([t.clone() for t in g["data"]], [t.clone() for t in g["target"]])
    xo, to = orc.regression_parse_input(g["data"], g["target"], g["L"], it)
    xg, tg = g["parsed"][it]
    for a in (xi, xo):
        assert torch.equal(a["model_input"], xg["model_input"]) and torch.equal(a["shock"], xg["shock"])
    assert torch.equal(ti, tg) and torch.equal(to, tg)
    # a batch without the shock field (exp 1 / 2 datasets): shock is None on every side
    xi, _ = pr.parse_input([t.clone() for t in g["data"][:4]], [t.clone() for t in g["target"]])
    assert xi["shock"] is None


def test_oracle_loss_functions_match_reference_incl_per_sample_and_masked():
    """oracle.elbo_loss / mvae_elbo_loss / *_per_sample against Reconstruction._elbo_loss / _mvae_elbo_loss
    of the live reference (problems.py:401-458): scalar and reduce=False forms, with and without a mask."""
    g = torch.load(os.path.join(GOLD, "losses.pt"), weights_only=False)
    gen = torch.Generator().manual_seed(g["seed"])
    B = g["B"]
    rv, rt = torch.randn(B, 3, 64, 64, generator=gen), torch.randn(B, 3, 64, 64, generator=gen)
    tv, tt = torch.rand(B, 3, 64, 64, generator=gen), torch.rand(B, 3, 64, 64, generator=gen)
    rp, tp = torch.rand(B, 7, generator=gen), torch.rand(B, 7, generator=gen)
    mu, lv = torch.randn(B, 256, generator=gen), 0.3 * torch.randn(B, 256, generator=gen)
    mask = (torch.rand(B, 3, 64, 64, generator=gen) > 0.5).float()
    klw, pm = g["kl_weight"], g["pose_multiplier"]
    got = {
        "elbo": orc.elbo_loss(rv, tv, mu, lv, klw),
        "elbo_masked": orc.elbo_loss(rv, tv, mu, lv, klw, mask),
        "elbo_ps": orc.elbo_per_sample(rv, tv, mu, lv, klw),
        "elbo_ps_masked": orc.elbo_per_sample(rv, tv, mu, lv, klw, mask),
        "mvae": orc.mvae_elbo_loss([rv, rt, rp], [tv, tt, tp], mu, lv, klw, pm),
        "mvae_masked": orc.mvae_elbo_loss([rv, rt], [tv, tt], mu, lv, klw, pm, mask),
        "mvae_ps": orc.mvae_elbo_per_sample([rv, rt, rp], [tv, tt, tp], mu, lv, klw, pm),
        "mvae_ps_masked": orc.mvae_elbo_per_sample([rv, rt], [tv, tt], mu, lv, klw, pm, mask),
    }
    for k, v in got.items():
        assert v.shape == g[k].shape, k
        assert torch.allclose(v, g[k], rtol=2e-6, atol=0), (k, v, g[k])
