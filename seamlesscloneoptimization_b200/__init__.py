"""B200-native seamlessClone(NORMAL_CLONE): hand-written sm_100a CUDA behind OpenCV's API.

The CUDA library is loaded lazily (first Context); importing the package needs no GPU.
"""
from ._capi import (EXEC_BLEND_PREFILLED, EXEC_DEFAULT, MEM_DEVICE, MEM_HOST, MIXED_CLONE, MONOCHROME_TRANSFER, NORMAL_CLONE, ScbError)
from .api import Context, Plan, SeamlessClone, default_context, seamlessClone

__all__ = [
    "Context", "Plan", "SeamlessClone", "seamlessClone", "default_context", "ScbError",
    "NORMAL_CLONE", "MIXED_CLONE", "MONOCHROME_TRANSFER", "MEM_HOST", "MEM_DEVICE", "EXEC_DEFAULT", "EXEC_BLEND_PREFILLED",
]
__version__ = "0.1.0"
