"""mmdyn_b200 — B200-native (sm_100a) implementation of the cnn-vae / cnn-mvae training and
inference step of SAIC-MONTREAL/multimodal-dynamics, behind the reference's `mmdyn.pytorch` API.

  mmdyn_b200.pytorch.models.models.setup_model   <- mmdyn/pytorch/models/models.py:13
  mmdyn_b200.pytorch.models.vae.{VAE,MVAE,...}   <- mmdyn/pytorch/models/vae.py
  mmdyn_b200.pytorch.problems.problems.*         <- mmdyn/pytorch/problems/problems.py
  mmdyn_b200.pytorch.main                        <- mmdyn/pytorch/main.py (same flags)

The math runs in libmmdyn_b200.so (hand-written CUDA, C ABI in include/mmdyn_b200.h).
"""
__version__ = "0.1.0"
