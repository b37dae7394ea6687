"""cusift_b200 — B200-native SIFT hot path behind cuSIFT's API.

The product is the CUDA shared library ``libcusift_b200.so`` (C ABI in
``include/cusift_b200.h``, reference-compatible C++ API in ``include/cusift/``).
This package is the Python harness around it (ctypes binding, synthetic inputs,
frame sharding); it contains no CPU implementation of the path.
"""
from ._lib import CsbParams, LIB_PATH, SIGNATURES, lib  # noqa: F401
from .api import COMPACT_DTYPE, SIFT_DTYPE, Context, CsbError, PinnedArray, align_up, make_params  # noqa: F401
from .sharding import shard_frames, shard_pairs, pair_index, all_pairs  # noqa: F401
from .synth import synth  # noqa: F401

__version__ = "0.1.0"
