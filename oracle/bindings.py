"""Bindings and build recipes of the CHECKER: oracle/liboracle.so (the FP64 restatement, rt_oracle.cpp) and
oracle/_ref/* (the reference's own headers / kernel.cu compiled here by build_ref.py).

Test infrastructure.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing in the product package does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from raytracinginoneweekendincuda_b200 import _abi as A  # noqa: E402  (the scene/camera structs the oracle consumes)


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd: list[str]) -> None:
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def oracle_path() -> str:
    return os.path.join(HERE, "liboracle.so")


def ref_stream_path(fast: bool = False) -> str:
    """libref_stream.so: the bit-exact pin (-O2 -ffp-contract=off); libref_stream_fast.so: the same sources built
    -O3 -march=native for the cpu_baseline timing only."""
    return os.path.join(HERE, "_ref", "libref_stream_fast.so" if fast else "libref_stream.so")


def ref_gpu_path() -> str:
    return os.path.join(HERE, "_ref", "ref_gpu")


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "rt_oracle.cpp")
    target = oracle_path()
    if force or not _newer(target, [src, os.path.join(ROOT, "include", "rt_abi.h"), os.path.join(ROOT, "include", "rt_rng.h")]):
        _run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-pthread", "-Wall", "-Wextra",
              src, "-o", target])
    return target


def build_ref() -> None:
    """oracle/_ref from /root/reference, when it is there (build container only)."""
    if not os.path.isdir(os.environ.get("RT_REFERENCE_DIR", "/root/reference")):
        return
    need = [ref_stream_path(), ref_stream_path(fast=True), ref_gpu_path()]
    srcs = [os.path.join(HERE, f) for f in ("build_ref.py", "ref_stream_main.cpp", "ref_image.cpp")]
    for base, _, files in os.walk(os.path.join(HERE, "shim")):
        srcs += [os.path.join(base, f) for f in files]
    if all(_newer(t, srcs) for t in need):
        return
    _run([sys.executable, os.path.join(HERE, "build_ref.py")])


def load_oracle() -> C.CDLL:
    if not os.path.exists(oracle_path()):
        build_oracle()
    lib = C.CDLL(oracle_path())
    declare_oracle(lib)
    return lib


def load_ref_stream(fast: bool = False):
    """-> CDLL or None when oracle/_ref was not built (it needs /root/reference)."""
    path = ref_stream_path(fast)
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    declare_ref_stream(lib)
    return lib


# oracle/rt_oracle.cpp
class oracle_stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("paths", C.c_uint64), ("box_tests", C.c_uint64),
                ("sphere_tests", C.c_uint64), ("quad_tests", C.c_uint64), ("medium_tests", C.c_uint64),
                ("draws", C.c_uint64), ("n_nodes", C.c_int32), ("n_objects", C.c_int32),
                ("medium_visits", C.c_int32 * 8)]


# oracle/ref_stream_main.cpp
class ref_stream_stats(C.Structure):
    _fields_ = [("rays", C.c_ulonglong), ("paths", C.c_ulonglong), ("draws", C.c_ulonglong),
                ("scene_draws", C.c_ulonglong), ("n_objects", C.c_int), ("n_nodes", C.c_int)]


def declare_oracle(lib: C.CDLL) -> None:
    lib.oracle_render.restype = C.c_int
    lib.oracle_render.argtypes = [C.POINTER(A.rt_scene_desc), C.POINTER(A.rt_camera), C.c_int, C.c_int, C.c_uint32,
                                  C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(oracle_stats)]
    lib.oracle_render_region.restype = C.c_int
    lib.oracle_render_region.argtypes = [C.POINTER(A.rt_scene_desc), C.POINTER(A.rt_camera), C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.POINTER(oracle_stats)]
    lib.oracle_trace_path.restype = C.c_int
    lib.oracle_trace_path.argtypes = [C.POINTER(A.rt_scene_desc), C.POINTER(A.rt_camera), C.c_int, C.c_int, C.c_int,
                                      C.c_uint32, C.c_void_p, C.c_int]
    lib.oracle_rng_uniform.restype = C.c_float
    lib.oracle_rng_uniform.argtypes = [C.c_uint32] * 6
    lib.oracle_bvh_topology.restype = C.c_int
    lib.oracle_bvh_topology.argtypes = [C.POINTER(A.rt_scene_desc), C.c_void_p, C.c_int32]
    lib.oracle_light_pdf.restype = C.c_int
    lib.oracle_light_pdf.argtypes = [C.POINTER(A.rt_scene_desc), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.oracle_light_direction.restype = C.c_int
    lib.oracle_light_direction.argtypes = [C.POINTER(A.rt_scene_desc), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.oracle_texture_value.restype = C.c_int
    lib.oracle_texture_value.argtypes = [C.POINTER(A.rt_scene_desc), C.c_int, C.c_double, C.c_double, C.c_void_p,
                                         C.c_void_p]


def declare_ref_stream(lib: C.CDLL) -> None:
    lib.ref_stream_render.restype = C.c_int
    lib.ref_stream_render.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(ref_stream_stats)]
    lib.ref_stream_scene_boxes.restype = C.c_int
    lib.ref_stream_scene_boxes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                           C.c_int, C.POINTER(C.c_ulonglong)]
    lib.ref_load_image_rgb8.restype = C.c_int
    lib.ref_load_image_rgb8.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_int]
