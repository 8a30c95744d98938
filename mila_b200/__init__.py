"""mila_b200 — B200-native (sm_100a) quantized Linear path behind Mila's kernel-launcher boundary.

Only what the hot path needs: csrc/ (CUDA kernels + C-ABI -> libmila_b200_linear.so), the ctypes
binding (_lib), the host-side mirror of Mila's Linear component (linear) and the tensor-parallel
sharding helpers (tp).  See DESIGN.md.
"""
from ._lib import InvalidArgument, LogicError, MilaB200Error, lib, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
