"""Shared fixtures.

`-m "not gpu"` tests run on a CPU-only box: they exercise the oracle, the host
scene surface and the library's exports.  `-m gpu` tests are the parity tests
proper and render through the C ABI on a B200.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import bindings as O  # noqa: E402  (the checker's bindings live with the checker)
from raytracinginoneweekendincuda_b200 import _abi as A  # noqa: E402
from raytracinginoneweekendincuda_b200 import _build  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the in-tree libraries exist (compiles them when a toolchain is here)."""
    cli = os.path.join(ROOT, "raytracinginoneweekendincuda_b200", "rt_cli")
    if not (os.path.exists(_build.lib_path()) and os.path.exists(cli)):
        _build.build_product()
    O.build_oracle()  # (a no-op unless rt_oracle.cpp or the ABI headers are newer than liboracle.so)
    return True


@pytest.fixture(scope="session")
def lib(built):
    from raytracinginoneweekendincuda_b200 import load_library
    return load_library()


@pytest.fixture(scope="session")
def oracle(built):
    return O.load_oracle()


@pytest.fixture(scope="session")
def ref_stream():
    """The reference's own classes compiled for the host (only where oracle/_ref was built)."""
    r = O.load_ref_stream()
    if r is None:
        pytest.skip("oracle/_ref/libref_stream.so not built (needs /root/reference)")
    return r


@pytest.fixture(scope="session")
def earth():
    """The earthmap texels as the reference's RtwImage produces them (fixture made by golden/make_golden.py)."""
    path = os.path.join(GOLDEN, "earthmap_rgb8.npz")
    assert os.path.exists(path), "tests/golden/earthmap_rgb8.npz missing"
    return np.ascontiguousarray(np.load(path)["rgb"])


def oracle_render(oracle, scene, cam, s0, s1, seed=1984, bvh=1, precision=64, threads=0):
    out = np.zeros((cam.image_height, cam.image_width, 3), np.float64)
    st = O.oracle_stats()
    rc = oracle.oracle_render(scene.desc, C.byref(cam), s0, s1, seed, bvh, precision, threads or (os.cpu_count() or 1),
                              out.ctypes.data, C.byref(st))
    assert rc == 0
    return out, st


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def oracle_render_region(oracle, scene, cam, x0, y0, w, h, s0, s1, seed=1984, bvh=1, threads=0):
    """Window [x0,x0+w) x [y0,y0+h) of the frame, streams keyed on the global pixel index: (sum[h,w,3], stats)."""
    out = np.zeros((h, w, 3), np.float64)
    st = O.oracle_stats()
    rc = oracle.oracle_render_region(scene.desc, C.byref(cam), x0, y0, w, h, s0, s1, seed, bvh, 64,
                                     threads or (os.cpu_count() or 1), out.ctypes.data, C.byref(st))
    assert rc == 0
    return out, st
