"""Model factory — mirror of mmdyn/pytorch/models/models.py:9-25 (`Regressor`, :28-77, belongs to
`--problem-type regression`, which is outside the accelerated path: SURVEY.md §8f)."""
from mmdyn_b200.pytorch import config
from mmdyn_b200.pytorch.models.vae import VAE, MVAE, Swish  # noqa: F401


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def setup_model(model_name, cross_modal=False, **kwargs):
    """Same dispatch rules as the reference: 'mvae' + cross-modal input -> MVAE, any other
    '...vae' -> VAE (which refuses cross-modal input)."""
    assert (model_name in config.MODELS), "Model is not implement yet"
    if 'mvae' in model_name and cross_modal:
        return MVAE(**kwargs)
    if 'vae' in model_name:
        assert not cross_modal, "VAE does not work with cross modal inputs."
        return VAE(**kwargs)
    if 'regressor' in model_name:
        raise NotImplementedError("the pose Regressor baseline is outside the B200 hot path (SURVEY.md §8f)")
    raise SystemExit("The model and modality combination is not valid.")
