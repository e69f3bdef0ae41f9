"""voxlogica_b200 -- B200-native execution backend for VoxLogicA's ImgQL operators.

The product is ``libvoxcuda.so`` (C ABI in ``include/voxcuda.h``, CUDA sources in
``voxlogica_b200/csrc``).  This package is the thin host side used in place of the F#
evaluator, which is not runnable here: ``ops`` mirrors the operator table one call per
operator, ``imgql`` is a minimal ImgQL driver so specs run unchanged.
"""
from . import _abi  # noqa: F401
from .ops import Context, Image, VoxError  # noqa: F401

__all__ = ["Context", "Image", "VoxError"]
__version__ = "0.1.0"
