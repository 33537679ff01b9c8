"""Runs the UNMODIFIED reference (SAIC-MONTREAL/multimodal-dynamics, `mmdyn.pytorch`) as the timed baseline of
bench.py — the reference's own modules through its own public step, none of this repo's code on the path.

    install()        copies /root/reference/mmdyn/{__init__.py, pytorch/**.py} into the git-ignored
                     baseline/_ref/ (called by __graft_entry__.build() in the build container; the copy
                     travels to the GPU box with the snapshot, /root/reference does not).  `pip install
                     --target baseline/_ref /root/reference` is not usable: the reference's setup.py declares
                     `py_modules=['mmdyn']` (a module that does not exist) and no packages, so pip installs
                     nothing importable; the package tree is copied file by file instead (recorded in DESIGN.md).
    load()           imports the reference from baseline/_ref (or /root/reference) with the two utility
                     modules that fail headless stubbed (`utils.training` runs `stty size` at
                     import, `utils.plots` needs pyquaternion / matplotlib) -> (problems, setup_model)
    time_step(...)   problems.py:150-155 on a dataset-free SeqModeling object: zero_grad,
                     `_evaluate_model`, backward, torch.optim.Adam.step — on the CPU (all host threads) or,
                     for the stock-PyTorch-eager comparison BASELINE.md §3 names, on a CUDA device.
"""
import os
import shutil
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference"


def install(verbose=True):
    """Copy the reference's python package (sources only, unmodified) to baseline/_ref/.  No-op when the
    reference tree is absent (GPU box) — then the copy made in the build container is used."""
    src = os.path.join(REF_SRC, "mmdyn")
    if not os.path.isdir(src):
        return os.path.isdir(os.path.join(REF_COPY, "mmdyn"))
    dst = os.path.join(REF_COPY, "mmdyn")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    shutil.copy2(os.path.join(src, "__init__.py"), os.path.join(dst, "__init__.py"))
    n = 0
    for dp, dn, fn in os.walk(os.path.join(src, "pytorch")):
        rel = os.path.relpath(dp, src)
        os.makedirs(os.path.join(dst, rel), exist_ok=True)
        for f in fn:
            if f.endswith(".py"):
                shutil.copy2(os.path.join(dp, f), os.path.join(dst, rel, f))
                n += 1
    if verbose:
        print(f"baseline/_ref: copied {n} reference source files (mmdyn.pytorch) from {REF_SRC}")
    return True


def available():
    return os.path.isdir(os.path.join(REF_COPY, "mmdyn", "pytorch")) or os.path.isdir(os.path.join(REF_SRC, "mmdyn", "pytorch"))


def load():
    root = REF_COPY if os.path.isdir(os.path.join(REF_COPY, "mmdyn", "pytorch")) else REF_SRC
    if root not in sys.path:
        sys.path.insert(0, root)
    tr = types.ModuleType("mmdyn.pytorch.utils.training")
    tr.progress_bar = lambda *a, **k: None
    tr.save_pkl = lambda *a, **k: None
    pl = types.ModuleType("mmdyn.pytorch.utils.plots")
    pl.plot_pose_tensorboard = pl.plot_single_pose_tensorboard = lambda *a, **k: None
    sys.modules["mmdyn.pytorch.utils.training"] = tr
    sys.modules["mmdyn.pytorch.utils.plots"] = pl
    from mmdyn.pytorch.models.models import setup_model
    from mmdyn.pytorch.problems import problems
    return problems, setup_model, root


def make_problem(problems, setup_model, device, use_pose=True, kl_weight=1.0 / 50, seed=0):
    """Dataset-free SeqModeling (SURVEY.md §8c recipe), cnn-mvae visuotactile."""
    pr = object.__new__(problems.SeqModeling)
    pr.parameters = {"model_name": "cnn-mvae", "input_type": "visuotactile", "use_pose": use_pose,
                     "mask_loss": False, "problem_type": "seq_modeling"}
    pr._kl_weight, pr._pose_multiplier, pr._conditional = kl_weight, 1000.0, False
    pr._device = torch.device(device)
    torch.manual_seed(seed)
    pr._model = setup_model("cnn-mvae", cross_modal=True, condition_dim=0, input_dim=4096, architecture="cnn",
                            conditional=False, categorical_conditions=False, latent_size=256, use_pose=use_pose)
    pr._model.to(pr._device)
    pr._model.train()
    return pr


def synth(B, seed, device):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g).to(device)
    x = {"model_input": [r(B, 3, 64, 64), r(B, 3, 64, 64)], "input_object_pose": [r(B, 7)], "shock": None}
    t = {"target_output": [r(B, 3, 64, 64), r(B, 3, 64, 64)], "target_object_pose": [r(B, 7)], "loss_mask": None}
    return x, t


def time_step(B, steps, warmup, device="cpu", threads=None):
    """-> (seconds per step list, final loss, source root).  CUDA: timed with events around each step."""
    problems, setup_model, root = load()
    if threads:
        torch.set_num_threads(threads)
    pr = make_problem(problems, setup_model, device)
    opt = torch.optim.Adam(pr._model.parameters(), lr=1e-3)  # problems.py:138
    x, t = synth(B, 1, device)
    cuda = torch.device(device).type == "cuda"
    times, loss = [], None
    for i in range(warmup + steps):
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad()
        outputs, loss = pr._evaluate_model(x, t)
        loss.backward()
        opt.step()
        lv = loss.item()  # problems.py:156: the reference reads the loss every step
        if cuda:
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, lv, root


if __name__ == "__main__":
    install()
    ts, lv, root = time_step(int(sys.argv[1]) if len(sys.argv) > 1 else 16, 2, 1, threads=os.cpu_count())
    print(f"reference from {root}: {ts} s/step, loss {lv:.3f}")
