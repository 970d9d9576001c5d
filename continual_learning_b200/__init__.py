"""continual_learning_b200 — B200 (sm_100a) kernels + host glue for the U-Net continual-learning step.

Public surface (mirrors the reference's Python surface for this path):
    UNet                      drop-in for models.unet.UNet (same constructor, state_dict keys)
    CrossEntropyDistillLoss   drop-in for nn.CrossEntropyLoss() (+ optional distillation)
    FusedAdam                 drop-in for torch.optim.Adam
    metrics                   drop-in for the reference `metrics` module (used half)
    TrainStep                 the fused, CUDA-graph-able step used by bench.py
"""
from . import _lib, metrics, ops, voc  # noqa: F401
from .loss import CrossEntropyDistillLoss
from .optim import FusedAdam
from .step import TrainStep
from .unet import UNet

__all__ = ["UNet", "CrossEntropyDistillLoss", "FusedAdam", "TrainStep", "metrics", "ops", "voc"]
