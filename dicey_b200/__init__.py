"""dicey_b200 -- B200-native FM-index primer matching (the `hunt` / `search` hot path of
gear-genomics/dicey).  The compute path is the CUDA library ``libdicey_b200.so`` built from
``csrc/``; this package is the thin host-side mirror of the reference's driver interface.
There is no CPU fallback: importing :mod:`dicey_b200.api` without the built library, or using it
without a CUDA device, raises.
"""
from .api import (DiceyB200Error, HuntParams, HuntResult, Index, library, library_path)  # noqa: F401

__all__ = ["DiceyB200Error", "HuntParams", "HuntResult", "Index", "library", "library_path"]
