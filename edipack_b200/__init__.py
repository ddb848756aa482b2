"""edipack_b200 -- B200-native Lanczos H x v engine for EDIpack's NORMAL mode.

The product is ``libedgpu.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/edgpu.h``); :mod:`edipack_b200.host` mirrors the reference's host interface on top of
it.  No CPU fallback exists: without the shared library / a B200 every call raises.
"""
from ._abi import ABI_SYMBOLS, LIB_PATH, EdgpuError, load  # noqa: F401
from .host import *  # noqa: F401,F403
