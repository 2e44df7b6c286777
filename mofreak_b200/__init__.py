"""mofreak_b200 — B200-native (sm_100a) MoFREAK descriptor extraction + bag-of-words assignment.

The product is native: mofreak_b200/csrc (hand-written CUDA behind the C ABI of include/mofreak_b200.h) and
mofreak_b200/host (C++ classes with the reference's MoFREAKUtilities / MotionInterchangePatterns /
BagOfWordsRepresentation call surface).  This package only carries the build recipe and a ctypes binding
used by the tests and the benchmark.
"""
from . import build as _build  # noqa: F401
from .capi import (Context, MofreakError, ROW_DTYPE, FRAME_GAP, SYMBOLS, LIB_PATH, lib, synth_clip, HostBuffer,  # noqa: F401
                   freak_scale_index)
