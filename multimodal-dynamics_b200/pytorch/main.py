"""CLI — same flags as the reference's `mmdyn/pytorch/main.py:13-54`.

    python -m mmdyn_b200.pytorch.main --problem-type seq_modeling --input-type visuotactile \\
        --model-name cnn-mvae --use-pose --dataset-path synthetic

`--dataset-path synthetic[:n_sequences[:seq_length]]` selects a dataset-shaped random stand-in
(the PyBullet/ShapeNetSem data is not redistributable); any other path is read by the reference's
own loader when the reference package is importable.
"""
import argparse
import os

from mmdyn_b200.pytorch import config


def build_parser():
    p = argparse.ArgumentParser(description='PyTorch Training (B200-native step)')
    # Problem
    p.add_argument('--problem-type', default='seq_modeling', type=str, help='Problem type (default: seq_modeling)')
    p.add_argument('--model-name', default='cnn-mvae', type=str, help='Model architecture name')
    p.add_argument('--input-type', default='visual', type=str,
                   help='The input modality (valid: visual, tactile, visuotactile)')
    p.add_argument('--use-pose', action='store_true', default=False,
                   help='Use pose as additional modality, only works for MVAE (default: False)')
    p.add_argument('--lr', default=0.001, type=float, help='learning rate (default: 0.001)')
    p.add_argument('--dataset-path', default='~/dataset', type=str, help='Absolute path to the dataset.')
    p.add_argument('--batchsize', default=128, type=int, help='Batchsize (default: 128)')
    p.add_argument('--criterion', default='crossentropy', type=str, help='Training loss (default: crossentropy)')
    p.add_argument('--optimizer', default='Adam', type=str, help='Adam or SGD (default: Adam)')
    p.add_argument('--num-epochs', default=100, type=int, help='Number of training epochs (default: 100)')
    p.add_argument('--mask-loss', action='store_true', default=False,
                   help='Mask the reconstruction loss to the object segment (default: False)')
    p.add_argument('--vis-pose', action='store_true', default=False, help='Visualize pose (default: False)')
    p.add_argument('--pose-multiplier', default=1000, type=float, help='Multiplier for pose loss (default: 1000)')
    # Misc
    p.add_argument('--save-name', default='run', type=str, help='Run name used for logs/checkpoints (default: run)')
    p.add_argument('--no-cuda', action='store_true', default=False, help='Do not use CUDA (unsupported here)')
    # VAE specific
    p.add_argument('--kl-weight', type=float, default=1.0, help='KL weight in the loss of VAE models (default: 1)')
    p.add_argument('--latent-size', type=int, default=256, help='Latent dimension (default: 256)')
    p.add_argument('--annealing-epochs', type=int, default=50, help='Number of epochs to anneal KL for (default: 50)')
    p.add_argument('--conditional', action='store_true', default=False, help='Conditional VAE: condition the image experts on the shock force (data[4])')
    # not in the reference (it can only write checkpoints, problems.py:751-757): continue a run from one
    p.add_argument('--resume', default=None, type=str,
                   help='epoch_N.ckpt written by this code or by the reference: load the model, continue at epoch N+1')
    return p


def main(argv=None):
    from mmdyn_b200.pytorch.problems.problems import Regression, Reconstruction, SeqModeling, DynModeling
    from mmdyn_b200.pytorch.utils.training import save_pkl
    args = build_parser().parse_args(argv)
    assert args.problem_type in config.PROBLEM_TYPES, "Invalid problem type."
    cls = {'regression': Regression, 'reconstruction': Reconstruction, 'dyn_modeling': DynModeling}.get(
        args.problem_type, SeqModeling)
    problem = cls(args)
    save_pkl(args, os.path.join(problem.log_dir, 'problem.pkl'))
    problem.train()
    return problem


if __name__ == "__main__":
    main()
