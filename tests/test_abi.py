"""The C-ABI library loads on a CPU-only box and exports what include/*.h declares."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, A, has_gpu


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)) - {"rt_scene_s", "rt_host_scene_s"})


@pytest.mark.parametrize("header", ["rt_abi.h", "rt_scenes_c.h"])
def test_every_declared_symbol_is_exported(lib, header):
    names = declared_functions(header)
    assert len(names) >= 6
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/{header} but not exported"


def test_struct_sizes_match_the_library(lib):
    for name in ["rt_prim", "rt_xform", "rt_object", "rt_material", "rt_texture", "rt_perlin", "rt_image",
                 "rt_scene_desc", "rt_camera", "rt_upload_options", "rt_render_params", "rt_stats", "rt_scene_info",
                 "rt_pack_info", "rt_timing"]:
        assert lib.rt_abi_sizeof(name.encode()) == C.sizeof(getattr(A, name)), name


def test_no_cpu_fallback_upload_fails_loudly_without_gpu(lib):
    """A valid scene on a GPU-less machine must be refused, not rendered some other way."""
    if has_gpu():
        pytest.skip("GPU present")
    from raytracinginoneweekendincuda_b200 import BuiltinScene
    sc = BuiltinScene(10)
    h = C.c_void_p()
    opt = A.rt_upload_options(device=0, bvh=A.RT_BVH_SAH)
    rc = lib.rt_scene_upload(sc.desc, C.byref(opt), C.byref(h))
    assert rc == A.RT_ERR_NO_DEVICE
    assert b"no CUDA device" in lib.rt_last_error()
    assert not h.value


def test_malformed_scenes_are_rejected(lib):
    from raytracinginoneweekendincuda_b200 import BuiltinScene
    sc = BuiltinScene(7)
    d = sc.desc.contents
    h = C.c_void_p()
    opt = A.rt_upload_options(device=0, bvh=A.RT_BVH_SAH)

    bad = A.rt_scene_desc.from_buffer_copy(d)
    bad.abi_version = 99
    assert lib.rt_scene_upload(C.byref(bad), C.byref(opt), C.byref(h)) == A.RT_ERR_INVALID

    prims = (A.rt_prim * d.n_prims)(*[d.prims[i] for i in range(d.n_prims)])
    prims[3].material = 1000
    bad = A.rt_scene_desc.from_buffer_copy(d)
    bad.prims = prims
    assert lib.rt_scene_upload(C.byref(bad), C.byref(opt), C.byref(h)) == A.RT_ERR_INVALID
    assert b"material" in lib.rt_last_error()

    bad = A.rt_scene_desc.from_buffer_copy(d)
    bad.n_objects = 0
    assert lib.rt_scene_upload(C.byref(bad), C.byref(opt), C.byref(h)) == A.RT_ERR_INVALID
    assert lib.rt_scene_upload(None, C.byref(opt), C.byref(h)) == A.RT_ERR_INVALID


def test_release_cached_memory_is_callable_without_a_device(lib):
    assert lib.rt_release_cached_memory() == A.RT_OK


def test_unknown_scene_id_is_an_error(lib):
    h = C.c_void_p()
    assert lib.rt_host_scene_builtin(77, None, 0, 0, C.byref(h)) == A.RT_ERR_INVALID
    assert b"unknown scene" in lib.rt_last_error()


def test_ppm_writers_match_the_reference_format(lib, tmp_path):
    """kernel.cu:696-723: 'P3\\nW H\\n255\\n' then 'r g b\\n' per pixel, top row first; P6 holds the same bytes.
    Host-only code: runs without a GPU."""
    import numpy as np
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    img[0, 0] = (0, 9, 255)
    p3, p6 = tmp_path / "a.ppm", tmp_path / "b.ppm"
    assert lib.rt_write_ppm(str(p3).encode(), img.ctypes.data, 53, 37) == A.RT_OK
    assert lib.rt_write_ppm_binary(str(p6).encode(), img.ctypes.data, 53, 37) == A.RT_OK
    want = "P3\n53 37\n255\n" + "".join("%d %d %d\n" % tuple(px) for px in img.reshape(-1, 3))
    assert p3.read_text() == want
    raw = p6.read_bytes()
    assert raw.startswith(b"P6\n53 37\n255\n") and raw[len(b"P6\n53 37\n255\n"):] == img.tobytes()
    assert lib.rt_write_ppm(b"/nonexistent-dir/x.ppm", img.ctypes.data, 53, 37) == A.RT_ERR_INVALID
    assert lib.rt_write_ppm(str(p3).encode(), None, 53, 37) == A.RT_ERR_INVALID
