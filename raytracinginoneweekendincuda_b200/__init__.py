"""B200-native render path for RayTracinginOneWeekendinCUDA scenes.

The product is ``librt_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/rt_abi.h``); this package is the thin host binding the tests, the
benchmark and Python callers use.  There is no CPU rendering path: importing
works without a GPU (host-side scene building does too), rendering needs one,
and a missing or stale library is an error, never a silent fallback.
"""
from __future__ import annotations

from ._lib import load_library, library_path  # noqa: F401
from .api import (BuiltinScene, Renderer, SCENE_NAMES, load_image, render_scene,  # noqa: F401
                  write_ppm)

__all__ = ["load_library", "library_path", "BuiltinScene", "Renderer", "SCENE_NAMES", "load_image",
           "render_scene", "write_ppm"]
