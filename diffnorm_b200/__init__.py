"""diffnorm_b200 — B200-native (sm_100a) implementation of DiffNorm's latent-diffusion normalization pass.

Importing the package loads the in-tree CUDA library (diffnorm_b200/csrc/libdiffnorm_b200.so); there is no
CPU or torch fallback — a missing library is an ImportError with build instructions.
"""
from . import _lib  # noqa: F401  (fails loudly when the extension is not built)
from .config import DiffNormConfig, UNIT_OFFSET, VOCAB, TIMESTEPS  # noqa: F401

__all__ = ["DiffNormConfig", "UNIT_OFFSET", "VOCAB", "TIMESTEPS"]
