"""pytorch_glow_b200 -- B200-native (sm_100a) Glow flow engine.

Drop-in for the flow hot path of corenel/pytorch-glow: the layer classes of
network/module.py and FlowStep / FlowModel / Glow of network/model.py, with the
same names, signatures and state_dict keys.  All arithmetic runs in hand-written
CUDA kernels behind the C ABI of include/glowk.h (libglowk.so); there is no CPU
or eager-PyTorch fallback.
"""
from . import config, snapshot
from .inferer import Inferer
from .model import FlowStep, FlowModel, FlowNet, Glow
from .module import (ActNorm, LinearZeros, Conv2d, Conv2dZeros, CouplingNet, f, Invertible1x1Conv,
                     Permutation2d, GaussianDiag, Split2d, Squeeze2d)

# names used by BASELINE.json's north_star for the same classes
InvertibleConv1x1 = Invertible1x1Conv
SqueezeLayer = Squeeze2d

__all__ = [
    "FlowStep", "FlowModel", "FlowNet", "Glow", "ActNorm", "LinearZeros", "Conv2d", "Conv2dZeros",
    "CouplingNet", "f", "Invertible1x1Conv", "InvertibleConv1x1", "Permutation2d", "GaussianDiag",
    "Split2d", "Squeeze2d", "SqueezeLayer", "config", "snapshot", "Inferer",
]
