"""mulskips_b200 -- B200-native KMC event loop of MulSKIPS behind a C ABI (include/mskc.h).

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI, built into lib/libmskc.so) and a thin
ctypes mirror of that ABI (capi.py).  There is no CPU fallback: importing works without a GPU (the
library loads), creating a context does not.
"""
from .capi import Mskc, MskcError, lib_path, load_library, ParityResult, ParallelResult  # noqa: F401

__all__ = ["Mskc", "MskcError", "lib_path", "load_library"]
