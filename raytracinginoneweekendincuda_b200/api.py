"""Host-side Python binding over the C ABI (include/rt_abi.h).

Mirrors the reference's host driver (reference kernel.cu:570-742): pick a
scene, set the image size / samples / depth, render, read the framebuffer,
write a PPM.  All rendering goes through rt_scene_upload / rt_render /
rt_readback; numpy only carries host buffers.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional, Sequence

import numpy as np

from . import _abi as A
from ._lib import load_library

SCENE_NAMES = {
    0: "bouncing_spheres", 1: "checkered_spheres", 2: "earth", 3: "perlin_spheres", 4: "quads",
    5: "simple_light", 6: "cornell_box", 7: "cornell_boxes", 8: "cornell_smoke", 9: "final_scene",
    10: "book1_final",
}



class RtError(RuntimeError):
    pass


def _check(lib, rc: int, what: str) -> None:
    if rc != A.RT_OK:
        raise RtError(f"{what} failed ({rc}): {lib.rt_last_error().decode(errors='replace')}")


def load_image(path: str) -> np.ndarray:
    """RtwImage::Load (reference RtwImage.h:51-87): decodes a baseline JPEG in the host library with stb_image's
    arithmetic and returns the texels ImageTexture samples -- linearised and re-quantised RGB8, [H, W, 3], row 0 = top."""
    lib = load_library()
    w, h = C.c_int32(), C.c_int32()
    _check(lib, lib.rt_image_load(path.encode(), C.byref(w), C.byref(h), None, 0), "rt_image_load")
    out = np.empty((h.value, w.value, 3), np.uint8)
    _check(lib, lib.rt_image_load(path.encode(), C.byref(w), C.byref(h), out.ctypes.data, out.size), "rt_image_load")
    return out


class BuiltinScene:
    """One of the reference's scenes (ids 0..9, kernel.cu:163-172; 10 = Book 1
    final) built on the host into the flat FP64 description."""

    def __init__(self, scene_id: int, earth=None):
        """`earth`: the image texture of scenes 2 and 9 -- a path to earthmap.jpg (decoded by the host library), or
        texels [H, W, 3] uint8 as RtwImage produces them, or None (cyan, Texture.h:113-114)."""
        self.lib = load_library()
        self.scene_id = scene_id
        self._earth = None
        ew = eh = 0
        ptr = None
        if isinstance(earth, (str, os.PathLike)):
            earth = load_image(os.fspath(earth))
        if earth is not None:
            self._earth = np.ascontiguousarray(earth, dtype=np.uint8)
            eh, ew = self._earth.shape[0], self._earth.shape[1]
            ptr = self._earth.ctypes.data
        self._h = C.c_void_p()
        _check(self.lib, self.lib.rt_host_scene_builtin(scene_id, ptr, ew, eh, C.byref(self._h)),
               "rt_host_scene_builtin")

    @property
    def desc(self):
        return self.lib.rt_host_scene_desc(self._h)

    def camera(self, width: int, height: int, spp: int, max_depth: int = 50) -> A.rt_camera:
        cam = A.rt_camera()
        _check(self.lib, self.lib.rt_host_scene_camera(self._h, width, height, spp, max_depth, C.byref(cam)),
               "rt_host_scene_camera")
        return cam

    @property
    def rng_draws(self) -> int:
        return int(self.lib.rt_host_scene_rng_draws(self._h))

    @property
    def reference_bvh_nodes(self) -> int:
        return int(self.lib.rt_host_scene_reference_bvh_nodes(self._h))

    def desc_bytes(self) -> int:
        d = self.desc.contents
        n = (d.n_objects * C.sizeof(A.rt_object) + d.n_prims * C.sizeof(A.rt_prim) +
             d.n_xforms * C.sizeof(A.rt_xform) + d.n_materials * C.sizeof(A.rt_material) +
             d.n_textures * C.sizeof(A.rt_texture) + d.n_perlins * C.sizeof(A.rt_perlin))
        for k in range(d.n_images):
            n += d.images[k].width * d.images[k].height * 3
        return n

    def close(self) -> None:
        if self._h:
            self.lib.rt_host_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    """A scene resident on one GPU -- or, with `devices=[...]`, replicated on several GPUs driven by this one
    process (rt_render splits the sample range over them, rt_readback reduces on the first): rt_scene_upload ..
    rt_scene_free."""

    def __init__(self, desc, device: int = 0, bvh: int = A.RT_BVH_SAH, max_leaf_prims: int = 0,
                 upload_flags: int = 0, devices: Optional[Sequence[int]] = None):
        self.lib = load_library()
        opt = A.rt_upload_options(device=device, bvh=bvh, max_leaf_prims=max_leaf_prims, flags=upload_flags)
        self._ids = None
        if devices is not None and len(devices) > 0:
            self._ids = (C.c_int32 * len(devices))(*devices)
            opt.n_devices = len(devices)
            opt.device_ids = self._ids
            device = devices[0]
        self._h = C.c_void_p()
        _check(self.lib, self.lib.rt_scene_upload(desc, C.byref(opt), C.byref(self._h)), "rt_scene_upload")
        self.device = device
        self._cam = None

    def render(self, cam: A.rt_camera, sample_begin: int = 0, sample_end: Optional[int] = None, seed: int = 1984,
               clear: bool = True, stream: int = 0, accum_ptr: int = 0, block_threads: int = 0,
               blocks_per_sm: int = 0, flags: int = 0, variant: int = A.RT_VARIANT_AUTO) -> None:
        if sample_end is None:
            sample_end = cam.samples_per_pixel
        p = A.rt_render_params(sample_begin=sample_begin, sample_end=sample_end, seed=seed, variant=variant,
                               clear=1 if clear else 0, block_threads=block_threads, blocks_per_sm=blocks_per_sm,
                               flags=flags, stream=stream or None, accum=accum_ptr or None)
        self._cam = cam
        _check(self.lib, self.lib.rt_render(self._h, C.byref(cam), C.byref(p)), "rt_render")

    def render_progressive(self, cam: A.rt_camera, batch: int, on_frame: Callable, sample_begin: int = 0,
                           sample_end: Optional[int] = None, seed: int = 1984, linear: bool = False,
                           srgb8: bool = True, variant: int = A.RT_VARIANT_AUTO) -> None:
        """Renders [sample_begin, sample_end) in batches; after each batch `on_frame(linear|None, srgb8|None,
        samples_done, samples_total)` gets the frame so far (numpy views of the handle's pinned buffers, valid
        during the call) while the next batch renders."""
        if sample_end is None:
            sample_end = cam.samples_per_pixel
        H, W = cam.image_height, cam.image_width

        def trampoline(_user, lin, s8, done, total):
            a = np.ctypeslib.as_array(lin, shape=(H, W, 3)) if lin else None
            b = np.ctypeslib.as_array(s8, shape=(H, W, 3)) if s8 else None
            on_frame(a, b, done, total)

        cb = A.rt_progress_fn(trampoline)
        p = A.rt_render_params(sample_begin=sample_begin, sample_end=sample_end, seed=seed, variant=variant, clear=1)
        self._cam = cam
        _check(self.lib, self.lib.rt_render_progressive(self._h, C.byref(cam), C.byref(p), batch, 1 if linear else 0,
                                                        1 if srgb8 else 0, cb, None), "rt_render_progressive")

    def timing(self) -> A.rt_timing:
        t = A.rt_timing()
        _check(self.lib, self.lib.rt_get_timing(self._h, C.byref(t)), "rt_get_timing")
        return t

    def sync(self) -> None:
        _check(self.lib, self.lib.rt_sync(self._h), "rt_sync")

    def accum_ptr(self):
        ptr = C.c_void_p()
        n = C.c_uint64()
        _check(self.lib, self.lib.rt_accum_ptr(self._h, C.byref(ptr), C.byref(n)), "rt_accum_ptr")
        return ptr.value, n.value

    def readback(self, linear: bool = True, srgb8: bool = False, accum_ptr: int = 0):
        """-> (linear float32 [H,W,3] row 0 = bottom | None, srgb8 uint8 [H,W,3] row 0 = top | None, rt_stats)"""
        cam = self._cam
        H, W = cam.image_height, cam.image_width
        lin = np.empty((H, W, 3), np.float32) if linear else None
        s8 = np.empty((H, W, 3), np.uint8) if srgb8 else None
        st = A.rt_stats()
        _check(self.lib, self.lib.rt_readback(self._h, accum_ptr or None, lin.ctypes.data if linear else None,
                                              s8.ctypes.data if srgb8 else None, C.byref(st)), "rt_readback")
        return lin, s8, st

    def info(self) -> A.rt_scene_info:
        i = A.rt_scene_info()
        _check(self.lib, self.lib.rt_scene_get_info(self._h, C.byref(i)), "rt_scene_get_info")
        return i

    def close(self) -> None:
        if self._h:
            self.lib.rt_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render_scene(scene_id: int, width: int, height: int, spp: int, max_depth: int = 50, seed: int = 1984,
                 device: int = 0, bvh: int = A.RT_BVH_SAH, earth=None, flags: int = 0):
    """Convenience: build, upload, render, read back.  Returns (linear, stats).  `earth`: see BuiltinScene."""
    sc = BuiltinScene(scene_id, earth)
    cam = sc.camera(width, height, spp, max_depth)
    r = Renderer(sc.desc, device=device, bvh=bvh)
    r.render(cam, 0, spp, seed=seed, flags=flags)
    lin, _, st = r.readback(linear=True)
    r.close()
    sc.close()
    return lin, st


def write_ppm(path: str, srgb8: np.ndarray, binary: bool = False) -> None:
    """The reference's P3 text PPM (kernel.cu:696-723), or binary P6; srgb8 is [H,W,3], top row first."""
    lib = load_library()
    a = np.ascontiguousarray(srgb8, dtype=np.uint8)
    fn = lib.rt_write_ppm_binary if binary else lib.rt_write_ppm
    _check(lib, fn(path.encode(), a.ctypes.data, a.shape[1], a.shape[0]), "rt_write_ppm")
