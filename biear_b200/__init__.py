"""biear_b200 -- B200-native (sm_100a) implementation of BiEAR's active-mode binaural front-end.

Only what the hot path needs lives here:
  csrc/        hand-written CUDA kernels + the C ABI (include/biear_b200.h) -> lib/libbiear_b200.so
  _lib.py      ctypes binding of that ABI (fails loudly when the library is missing)
  ops.py       torch tensors <-> C ABI (pointers + current stream), autograd nodes
  frontend.py  host-side mirror of the reference's front-end modules (model_torch.py:70-776)
  model_torch.py  drop-in namespace for `from model_torch import build_model_active, ...`
"""
from .frontend import (AuralNetGammatoneFB, BinauralAdaptiveGammatoneFB,  # noqa: F401
                       BinauralAdaptiveGammatoneFB_SingleController, FramewiseAdaptiveGammatoneFB,
                       FramewiseFixedGammatoneFB)
from .ops import cc_feature  # noqa: F401

__all__ = [
    "AuralNetGammatoneFB", "BinauralAdaptiveGammatoneFB", "BinauralAdaptiveGammatoneFB_SingleController",
    "FramewiseAdaptiveGammatoneFB", "FramewiseFixedGammatoneFB", "cc_feature",
]
