"""kimera-rpgo_b200 — B200-native PCM outlier-rejection hot path behind the Kimera-RPGO API.

The package is a thin host layer over an in-tree CUDA shared library (csrc/, C ABI in
include/rpgo_b200.h).  Importing it never touches the CPU oracle; using it without the built
library or without a GPU raises."""
from . import _capi  # noqa: F401
from .pcm import (BETWEEN, CLIQUE_EXACT, CLIQUE_HEU, CLIQUE_HEU_INCREMENTAL, KERNEL_AUTO, KERNEL_DIRECT,  # noqa: F401
                  KERNEL_TILED, MODE_PCM, MODE_SIMPLE, OTHER, PRIOR, TRAJ_FOLD, TRAJ_SCAN, PcmGpu, RpgoError)

__all__ = ["PcmGpu", "RpgoError"]
