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


class HandScene:
    """A scene description written by hand through the ctypes structs (what a non-C++ host would do): one textured
    sphere of radius 1, optionally under RotateY(degrees) then Translate(offset) -- the reference's
    `new Translate(new RotateY(sphere, degrees), offset)`."""

    def __init__(self, earth, degrees=0.0, offset=(0.0, 0.0, 0.0), checker=False):
        self.earth = np.ascontiguousarray(earth)
        self.prims = (A.rt_prim * 1)()
        p = self.prims[0]
        p.type = A.RT_PRIM_SPHERE
        p.material = 0
        p.radius = 1.0
        p.a[:] = [0.0, 0.0, 0.0]
        self.xforms = (A.rt_xform * 2)()
        n_x = 0
        if degrees != 0.0 or any(offset):
            self.xforms[0].type = A.RT_XFORM_TRANSLATE  # outermost first
            self.xforms[0].v[:] = list(offset)
            self.xforms[1].type = A.RT_XFORM_ROTATE_Y
            rad = np.radians(degrees)
            self.xforms[1].v[:] = [float(np.sin(rad)), float(np.cos(rad)), degrees]
            n_x = 2
            p.first_xform, p.xform_count = 0, 2
        self.objects = (A.rt_object * 1)()
        o = self.objects[0]
        o.kind = A.RT_OBJ_PRIM
        o.first_prim, o.prim_count = 0, 1
        o.bbox[:] = [offset[0] - 1.5, offset[0] + 1.5, offset[1] - 1.0, offset[1] + 1.0, offset[2] - 1.5, offset[2] + 1.5]
        self.materials = (A.rt_material * 1)()
        self.materials[0].type = A.RT_MAT_LAMBERTIAN
        self.materials[0].texture = 2 if checker else 0
        self.textures = (A.rt_texture * 3)()
        self.textures[0].type = A.RT_TEX_IMAGE
        self.textures[0].image = 0
        self.textures[1].type = A.RT_TEX_SOLID
        self.textures[1].color[:] = [0.9, 0.1, 0.1]
        self.textures[2].type = A.RT_TEX_CHECKER  # even = the image, odd = red: (u,v) reached through a checker
        self.textures[2].even, self.textures[2].odd, self.textures[2].scale = 0, 1, 0.7
        self.images = (A.rt_image * 1)()
        self.images[0].width, self.images[0].height = self.earth.shape[1], self.earth.shape[0]
        self.images[0].rgb = self.earth.ctypes.data_as(C.POINTER(C.c_uint8))
        self._desc = A.rt_scene_desc(abi_version=A.RT_ABI_VERSION, n_objects=1, n_prims=1, n_xforms=n_x, n_materials=1,
                                     n_textures=3, n_images=1, objects=self.objects, prims=self.prims, xforms=self.xforms,
                                     materials=self.materials, textures=self.textures, images=self.images)
        self.desc = C.pointer(self._desc)

    def camera(self, W, H, spp, max_depth=50):
        cam = A.rt_camera(image_width=W, image_height=H, samples_per_pixel=spp, max_depth=max_depth, vfov=40.0,
                          defocus_angle=0.0, focus_dist=1.0, aperture=0.0, time0=0.0, time1=0.0)
        cam.lookfrom[:] = [0.3, 0.8, 4.0]
        cam.lookat[:] = [0.3, 0.0, 0.0]
        cam.vup[:] = [0.0, 1.0, 0.0]
        cam.background[:] = [0.8, 0.9, 1.0]
        return cam


class BoxScene:
    """A scene of MakeBox lists written through the ctypes structs (reference Instance.h:166-184 for the six quads and
    their order): a GLASS box under RotateY + Translate (rays enter it, travel inside and leave through the exit face), a
    fuzzy METAL box, a ConstantMedium whose boundary is a rotated box, on a Lambertian ground sphere under a bright sky.
    Exercises the slab-tested box (csrc/rt_trace.cuh HitBox / BoxSpan) where the built-in scenes do not: hits from inside,
    two-sided faces, a medium entered from any side."""

    @staticmethod
    def _box_quads(lo, hi):
        dx, dy, dz = (hi[0] - lo[0], 0.0, 0.0), (0.0, hi[1] - lo[1], 0.0), (0.0, 0.0, hi[2] - lo[2])

        def neg(v):
            return tuple(-c for c in v)

        return [((lo[0], lo[1], hi[2]), dx, dy), ((hi[0], lo[1], hi[2]), neg(dz), dy), ((hi[0], lo[1], lo[2]), neg(dx), dy),
                ((lo[0], lo[1], lo[2]), dz, dy), ((lo[0], hi[1], hi[2]), dx, neg(dz)), ((lo[0], lo[1], lo[2]), dx, dz)]

    def __init__(self):
        boxes = [  # (lo, hi, degrees, offset, material, medium?)
            ((-1.0, 0.0, -1.0), (1.0, 2.0, 1.0), 25.0, (0.2, 0.0, 0.3), 1, False),    # glass
            ((-0.5, 0.0, -0.5), (0.5, 1.5, 0.5), 0.0, (3.0, 0.0, 0.0), 2, False),     # metal, axis-aligned, no instance
            ((-0.8, 0.0, -1.0), (0.8, 1.5, 1.0), -15.0, (-3.0, 0.0, 0.2), 3, True),   # smoke
        ]
        n_prims = 1 + 6 * len(boxes)
        self.prims = (A.rt_prim * n_prims)()
        self.xforms = (A.rt_xform * (2 * len(boxes)))()
        self.objects = (A.rt_object * (1 + len(boxes)))()
        g = self.prims[0]
        g.type, g.material, g.radius = A.RT_PRIM_SPHERE, 0, 1000.0
        g.a[:] = [0.0, -1000.0, 0.0]
        o = self.objects[0]
        o.kind, o.first_prim, o.prim_count = A.RT_OBJ_PRIM, 0, 1
        o.bbox[:] = [-1000, 1000, -2000, 0, -1000, 1000]
        n_x = 0
        for b, (lo, hi, deg, off, mat, medium) in enumerate(boxes):
            first_x, count_x = 0, 0
            rad = np.radians(deg)
            s, c = float(np.sin(rad)), float(np.cos(rad))
            if deg != 0.0:
                self.xforms[n_x].type = A.RT_XFORM_TRANSLATE  # outermost first: Translate(RotateY(box))
                self.xforms[n_x].v[:] = list(off)
                self.xforms[n_x + 1].type = A.RT_XFORM_ROTATE_Y
                self.xforms[n_x + 1].v[:] = [s, c, deg]
                first_x, count_x = n_x, 2
                n_x += 2
            corners = []
            for k, (q, u, v) in enumerate(self._box_quads(lo, hi)):
                p = self.prims[1 + 6 * b + k]
                p.type, p.material = A.RT_PRIM_QUAD, mat
                if deg == 0.0:
                    q = tuple(q[a] + off[a] for a in range(3))  # no instance: the box is built where it stands
                p.a[:] = list(q)
                p.b[:] = list(u)
                p.c[:] = list(v)
                p.first_xform, p.xform_count = first_x, count_x
                for i in (0, 1):
                    for j in (0, 1):
                        pt = np.array(q) + i * np.array(u) + j * np.array(v)
                        if deg != 0.0:  # Instance.h:136-147 then the offset
                            pt = np.array([c * pt[0] + s * pt[2], pt[1], -s * pt[0] + c * pt[2]]) + np.array(off)
                        corners.append(pt)
            corners = np.array(corners)
            ob = self.objects[1 + b]
            ob.kind = A.RT_OBJ_MEDIUM if medium else A.RT_OBJ_LIST
            ob.first_prim, ob.prim_count = 1 + 6 * b, 6
            if medium:
                ob.phase_material, ob.medium_id, ob.density = 4, 0, 0.8
            lo_w, hi_w = corners.min(0), corners.max(0)
            ob.bbox[:] = [lo_w[0], hi_w[0], lo_w[1], hi_w[1], lo_w[2], hi_w[2]]
        self.n_xforms = n_x
        self.materials = (A.rt_material * 5)()
        self.textures = (A.rt_texture * 3)()
        for t, col in enumerate([(0.5, 0.5, 0.5), (0.1, 0.1, 0.1), (0.9, 0.9, 0.9)]):
            self.textures[t].type = A.RT_TEX_SOLID
            self.textures[t].color[:] = list(col)
        self.materials[0].type, self.materials[0].texture = A.RT_MAT_LAMBERTIAN, 0
        self.materials[1].type, self.materials[1].ior = A.RT_MAT_DIELECTRIC, 1.5
        self.materials[2].type, self.materials[2].fuzz = A.RT_MAT_METAL, 0.1
        self.materials[2].albedo[:] = [0.8, 0.6, 0.2]
        self.materials[3].type, self.materials[3].texture = A.RT_MAT_LAMBERTIAN, 1  # (boundary quads: never shaded)
        self.materials[4].type, self.materials[4].texture = A.RT_MAT_ISOTROPIC, 2
        self._desc = A.rt_scene_desc(abi_version=A.RT_ABI_VERSION, n_objects=1 + len(boxes), n_prims=n_prims,
                                     n_xforms=n_x, n_materials=5, n_textures=3, objects=self.objects, prims=self.prims,
                                     xforms=self.xforms, materials=self.materials, textures=self.textures)
        self.desc = C.pointer(self._desc)

    def camera(self, W, H, spp, max_depth=50):
        cam = A.rt_camera(image_width=W, image_height=H, samples_per_pixel=spp, max_depth=max_depth, vfov=35.0,
                          defocus_angle=0.0, focus_dist=10.0, aperture=0.0, time0=0.0, time1=0.0)
        cam.lookfrom[:] = [0.5, 3.0, 10.0]
        cam.lookat[:] = [0.0, 1.0, 0.0]
        cam.vup[:] = [0.0, 1.0, 0.0]
        cam.background[:] = [0.7, 0.8, 1.0]
        return cam
