"""Module-name shim for the reference's Python binding.

The reference's test script does `from SeamlessClone import SeamlessClone`
(/root/reference/seamlessClone-CUDA/seamlessClone-python-binding/SeamlessClone_test.py:2, the Boost.Python
module built from SeamlessClone.cpp).  Put this directory on sys.path and that import resolves to the
B200 implementation with the same class and method names.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from seamlesscloneoptimization_b200.api import SeamlessClone  # noqa: E402,F401

__all__ = ["SeamlessClone"]
