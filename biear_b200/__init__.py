"""biear_b200 -- B200-native (sm_100a) implementation of BiEAR's active-mode binaural front-end.

Only what the hot path needs lives here:
  csrc/        hand-written CUDA kernels + the C ABI (include/biear_b200.h) -> lib/libbiear_b200.so
  _lib.py      ctypes binding of that ABI (fails loudly when the library is missing)
  ops.py       torch tensors <-> C ABI (pointers + current stream), autograd nodes
  graph.py     CUDA-graph capture of a whole step (forward + backward)
  dist.py      batch data-parallel plumbing (shards + one flat-bucket gradient all-reduce)
  frontend.py  host-side mirror of the reference's front-end modules (model_torch.py:70-776)
  model_torch.py  drop-in namespace for `from model_torch import build_model_active, ...`
"""
from .frontend import (AuralNetGammatoneFB, BinauralAdaptiveGammatoneFB,  # noqa: F401
                       BinauralAdaptiveGammatoneFB_SingleController, FramewiseAdaptiveGammatoneFB,
                       FramewiseFixedGammatoneFB)
from .graph import GraphedStep  # noqa: F401
from .ops import cc_feature, q_regularizers  # noqa: F401

__all__ = [
    "AuralNetGammatoneFB", "BinauralAdaptiveGammatoneFB", "BinauralAdaptiveGammatoneFB_SingleController",
    "FramewiseAdaptiveGammatoneFB", "FramewiseFixedGammatoneFB", "GraphedStep", "cc_feature", "q_regularizers",
]
