"""thevc_b200 -- B200-native (sm_100a) implementation of the HM-7.2 HEVC hot path (fr34k8/thevc).

Layout: ``csrc/`` CUDA kernels + the C ABI (``include/thevc_cuda.h``), ``host/`` the C++ mirror of
the reference's TComRdCost / TComInterpolationFilter / TComTrQuant / TComPrediction / TEncSearch
interfaces over that ABI, ``capi.py`` the ctypes binding used by the tests and the bench, and
``tlibcuda.py`` a thin object wrapper (pictures as numpy arrays in, results out).

The CUDA library is the only compute path.  Nothing here imports ``oracle/``.
"""
from .capi import load, lib_path  # noqa: F401
from .tlibcuda import TLibCuda, TvcError, HostPic  # noqa: F401

__all__ = ["load", "lib_path", "TLibCuda", "TvcError", "HostPic"]
