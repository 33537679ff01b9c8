"""The reference-facing entry points end to end on a B200: `main.py` flags -> Problem classes ->
fused step -> fused optimizer -> checkpoint, on the synthetic dataset stand-in (the PyBullet data
cannot be shipped).  Covers BASELINE.json configs [1]-[3] plus the dyn_modeling and --mask-loss
variants at tiny sizes."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def run_main(tmp_path, monkeypatch, extra):
    from mmdyn_b200.pytorch.main import main
    monkeypatch.chdir(tmp_path)
    argv = ["--dataset-path", "synthetic:8:5", "--batchsize", "4", "--num-epochs", "2", "--annealing-epochs", "4",
            "--save-name", "t"] + extra
    return main(argv)


def check_run(problem, n_keys):
    ck = sorted(glob.glob(os.path.join(problem.checkpoint_dir, "epoch_*.ckpt")))
    assert ck, "no best-loss checkpoint written"
    state = torch.load(ck[-1], weights_only=False)
    assert set(state) == {"model", "loss", "epoch"} and len(state["model"]) == n_keys
    assert all(torch.isfinite(v).all() for v in state["model"].values() if v.is_floating_point())
    losses = problem._logger_dict["Loss/train_epoch"]
    assert len(losses) == 2 and all(l == l and l < 1e7 for l in losses)
    assert problem._logger_dict["KL_annealing/train_epoch"] == [0.25, 0.5]
    assert os.path.exists(os.path.join(problem.log_dir, "results.pkl"))
    assert os.path.exists(os.path.join(problem.log_dir, "problem.pkl"))


def test_main_seq_modeling_mvae_visuotactile_pose(tmp_path, monkeypatch):
    p = run_main(tmp_path, monkeypatch, ["--problem-type", "seq_modeling", "--input-type", "visuotactile",
                                         "--model-name", "cnn-mvae", "--use-pose"])
    check_run(p, 106)
    for k in ("visual", "tactile", "pose"):
        assert len(p._logger_dict["Perf_measure_train/" + k]) == 2
    # the reference checkpoint format loads back into a fresh mirror model (and would into the reference)
    from mmdyn_b200.pytorch.models.models import setup_model
    m = setup_model("cnn-mvae", cross_modal=True, condition_dim=0, input_dim=4096, architecture="cnn",
                    conditional=False, categorical_conditions=False, latent_size=256, use_pose=True)
    ck = sorted(glob.glob(os.path.join(p.checkpoint_dir, "epoch_*.ckpt")))[-1]
    m.load_state_dict(torch.load(ck, weights_only=False)["model"])


def test_main_seq_modeling_vae_tactile(tmp_path, monkeypatch):
    p = run_main(tmp_path, monkeypatch, ["--problem-type", "seq_modeling", "--input-type", "tactile",
                                         "--model-name", "cnn-vae"])
    check_run(p, 46)
    assert len(p._logger_dict["Perf_measure_train/tactile"]) == 2


def test_main_dyn_modeling_mvae_masked(tmp_path, monkeypatch):
    # dyn_modeling feeds all S*L frames of the batch (4 sequences x 5 frames = 20 rows per step)
    p = run_main(tmp_path, monkeypatch, ["--problem-type", "dyn_modeling", "--input-type", "visuotactile",
                                         "--model-name", "cnn-mvae", "--mask-loss", "--optimizer", "SGD"])
    check_run(p, 92)


def run_main_cond(tmp_path, monkeypatch, extra):
    from mmdyn_b200.pytorch.main import main
    monkeypatch.chdir(tmp_path)
    return main(["--dataset-path", "synthetic:8:5:3", "--batchsize", "4", "--num-epochs", "2", "--annealing-epochs", "4",
                 "--save-name", "t", "--conditional"] + extra)


def test_main_conditional_mvae_pose_and_vae(tmp_path, monkeypatch):
    """--conditional (exp 3: shock force as the CVAE condition, data[4]) through main.py; the step runs as a
    CUDA graph with a static condition buffer."""
    p = run_main_cond(tmp_path, monkeypatch, ["--problem-type", "seq_modeling", "--input-type", "visuotactile",
                                              "--model-name", "cnn-mvae", "--use-pose"])
    check_run(p, 106)
    assert p._condition_dim == 3 and p._conditional
    sd = p._model.state_dict()
    assert tuple(sd["visual_encoder.linear_means.weight"].shape) == (256, 515)
    assert tuple(sd["tactile_decoder.upsample.0.weight"].shape) == (6400, 259)
    assert tuple(sd["pose_encoder.linear_means.weight"].shape) == (256, 512)  # the pose expert is un-conditional
    assert getattr(p, "_graph_cache", None), "the conditional step did not go through the CUDA graph"
    p = run_main_cond(tmp_path, monkeypatch, ["--problem-type", "dyn_modeling", "--input-type", "visual",
                                              "--model-name", "cnn-vae"])
    check_run(p, 46)


def test_main_regression(tmp_path, monkeypatch):
    """--problem-type regression --model-name regressor: first frame -> resting pose."""
    p = run_main(tmp_path, monkeypatch, ["--problem-type", "regression", "--model-name", "regressor",
                                         "--input-type", "tactile"])
    ck = sorted(glob.glob(os.path.join(p.checkpoint_dir, "epoch_*.ckpt")))
    state = torch.load(ck[-1], weights_only=False)
    assert set(state) == {"model", "loss", "epoch"} and len(state["model"]) == 27
    losses = p._logger_dict["Loss/train_epoch"]
    assert len(losses) == 2 and all(float(l) == float(l) for l in losses)
    assert len(p._logger_dict["Perf_measure_train/pose"]) == 2
    vl = [float(v) for v in p._logger_dict["Loss/validation_epoch"]]
    assert len(vl) == 2 and all(v == v and v < 1e4 for v in vl)
    p = run_main_cond(tmp_path, monkeypatch, ["--problem-type", "regression", "--model-name", "regressor",
                                              "--input-type", "visual"])
    assert tuple(p._model.state_dict()["out_net.0.weight"].shape) == (256, 515)


def test_resume_from_checkpoint(tmp_path, monkeypatch):
    """--resume (not in the reference, which only writes checkpoints): weights, BatchNorm buffers, best loss,
    epoch counter and the fused optimizer's moments / step count come back; training continues at epoch N+1."""
    from mmdyn_b200.pytorch.main import main
    p = run_main(tmp_path, monkeypatch, ["--problem-type", "seq_modeling", "--input-type", "visuotactile",
                                         "--model-name", "cnn-mvae"])
    ck = sorted(glob.glob(os.path.join(p.checkpoint_dir, "epoch_*.ckpt")))[-1]
    state = torch.load(ck, weights_only=False)
    assert os.path.exists(ck + ".optim")
    ost = torch.load(ck + ".optim", weights_only=False)
    steps_per_epoch = len(p.train_loader)
    assert ost["kind"] == "FusedAdam" and ost["step"] == steps_per_epoch * (state["epoch"] + 1)
    monkeypatch.chdir(tmp_path)
    argv = ["--dataset-path", "synthetic:8:5", "--batchsize", "4", "--annealing-epochs", "4", "--save-name", "r",
            "--problem-type", "seq_modeling", "--input-type", "visuotactile", "--model-name", "cnn-mvae", "--resume", ck]
    from mmdyn_b200.pytorch.problems.problems import SeqModeling
    from mmdyn_b200.pytorch.main import build_parser
    q = SeqModeling(build_parser().parse_args(argv + ["--num-epochs", str(state["epoch"] + 2)]))
    assert q._start_epoch == state["epoch"] + 1 and float(q._best_loss) == float(state["loss"])
    for k, v in q._model.state_dict().items():
        assert torch.equal(v.cpu(), state["model"][k].cpu()), k
    assert int(q._optimizer._step_dev.item()) == ost["step"]
    assert torch.equal(q._optimizer._bufs[0].cpu(), ost["bufs"][0])
    q.train()
    assert len(q._logger_dict["Loss/train_epoch"]) == 1  # exactly the one remaining epoch ran
    assert q._logger_dict["KL_annealing/train_epoch"] == [(state["epoch"] + 2) / 4]
    assert int(q._optimizer.state_dict()["step"]) == ost["step"] + steps_per_epoch
    # a reference-format checkpoint without the side file resumes too (optimizer restarts)
    os.remove(ck + ".optim")
    r = SeqModeling(build_parser().parse_args(argv + ["--num-epochs", "1"]))
    assert r._start_epoch == state["epoch"] + 1 and int(r._optimizer.state_dict()["step"]) == 0
    with pytest.raises(ValueError):
        bad = os.path.join(tmp_path, "bad.ckpt")
        torch.save({"model": {}}, bad)
        r.load_checkpoint(bad)


def test_unsupported_paths_fail_loudly(tmp_path, monkeypatch):
    with pytest.raises(RuntimeError, match="CUDA"):
        run_main(tmp_path, monkeypatch, ["--no-cuda"])
