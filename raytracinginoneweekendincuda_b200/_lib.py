"""Loads librt_b200.so.  Fails loudly: there is no fallback implementation."""
from __future__ import annotations

import ctypes as C
import os

from . import _abi

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path() -> str:
    # RT_B200_LIBRARY: development override (e.g. a build with other compiler flags)
    return os.environ.get("RT_B200_LIBRARY") or os.path.join(_PKG, "librt_b200.so")


def load_library() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the render path.")
    lib = C.CDLL(path)
    _abi.declare_host(lib)
    _abi.declare_device(lib)
    _LIB = lib
    return lib
